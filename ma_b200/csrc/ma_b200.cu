// libma_b200.so — C ABI (include/ma_b200.h) over the sm_100a kernels. Host side of the drop-in boundary.
#include "common.cuh"
#include "ksw.cuh"
#include <chrono>
#include "pipeline.cuh"
#include "index_build.cuh"
#include <algorithm>
#include <nvtx3/nvToolsExt.h> // header-only (the tools library is looked up at run time): ranges for nsys / ncu --nvtx
#include <numeric>
#include <condition_variable>
#include <mutex>
#include <stdexcept>
#include <thread>

using namespace ma;

struct KswHostBin
{
    int W; // window class of ksw_batch_kernel; 0 for the bins of ksw_qs_kernel
    std::vector<int> order;
    long long tb_stride = 0;
    int cig_stride = 0;
    int qs = 0; // ksw_qs_kernel: 1 + (blocks - 1) * 2 + right-aligned
};

// Tracing hook: one NVTX range per API call and pipeline stage; free when no tool is attached.
struct NvtxRange
{
    explicit NvtxRange( const char* name )
    {
        nvtxRangePushA( name );
    }
    ~NvtxRange( )
    {
        nvtxRangePop( );
    }
    NvtxRange( const NvtxRange& ) = delete;
};

struct ma_b200_ctx
{
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    ma_b200_params params;
    int64_t launches = 0;
    EventTimer timer;

    // ---- DP batch state
    int64_t ksw_n = 0;
    DevBuf<KswTask> ksw_tasks;
    DevBuf<unsigned char> ksw_seq;
    DevBuf<KswOut> ksw_out;
    DevBuf<unsigned int> ksw_cigar;
    long long ksw_cigar_cap = 0, ksw_cigar_bound = 0;
    DevBuf<unsigned char> ksw_tb;
    DevBuf<unsigned int> ksw_cigscratch;
    DevBuf<int> ksw_order;
    DevBuf<unsigned long long> ksw_ctrl; // [0] cigar cursor, [16] next (as int), [32] error (as int), [48] cells: one 128-byte line each
    std::vector<KswHostBin> ksw_bins;
    unsigned long long ksw_cigar_used = 0;
    std::vector<KswTask> ksw_host_tasks; // host copy (device tags): bins the problems ksw_qs_kernel hands over
    DevBuf<int> ksw_redo; // [0] count, [1 ..] task ids handed over by ksw_qs_kernel (standalone batches)
    cudaEvent_t binEv[ MA_NBINS ][ 2 ] = { { nullptr } }; // MA_B200_DP_BINS=1: per-launch times

    // ---- index (replicated per context / GPU)
    bool have_index = false;
    DevIndex index;
    DevBuf<U4> ix_bwt; // reference layout (what index_upload got / index_download returns)
    DevBuf<U4> ix_bwtp; // bit-plane layout the kernels read (fmindex.cuh relayout_block)
    DevBuf<long long> ix_sa;
    DevBuf<unsigned char> ix_pac;
    DevBuf<long long> ix_contigs; // starts then lengths
    int64_t ix_words = 0, ix_nsa = 0, ix_npac = 0;

    // ---- alignment batch state
    int64_t n_reads = 0, reads_bytes = 0;
    int max_read_len = 0;
    int stage_done = 0;
    bool ksw_extension_only = false; // ma_b200_ksw_set_extension_only
    int64_t batch_split = 0; // ma_b200_align_batch: reads per sub-batch of the pipelined form, 0 = one shot
    ma_b200_ctx* shadow = nullptr; // second set of slabs + stream for the pipelined ma_b200_align_batch
    DevBuf<unsigned char> reads;
    DevBuf<long long> read_off;
    DevBuf<ReadInfo> info;
    DevBuf<DSeed> seeds;
    DevBuf<SegRec> lists;
    DevBuf<FSeg> fsegs;
    DevBuf<SegRec> dbg_segs;
    DevBuf<int> dbg_nsegs;
    int dbg_cap = 0;
    DevBuf<unsigned char> harm_scratch;
    DevBuf<DSeed> set_seeds;
    DevBuf<SetHeader> sets;
    DevBuf<KswTask> tasks;
    DevBuf<int> bin_order;
    DevBuf<KswOut> task_out;
    DevBuf<unsigned int> task_cigar;
    DevBuf<DAln> alns;
    DevBuf<unsigned int> runs;
    DevBuf<unsigned int> run_scratch;
    DevBuf<PipeCtrl> ctrl;
    DevBuf<long long> pair_sc;
    DevBuf<int> pair_meta;
    PipeCtrl hctrl;
    int64_t n_seeds = 0, n_sets = 0, n_set_seeds = 0, n_tasks = 0, n_runs = 0, n_task_cigar = 0;
    cudaEvent_t ev[ 8 ] = { nullptr };
    // ---- ma_b200_align_batch, one-shot form: copies on a second stream under the kernels
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_copy[ 3 ] = { nullptr }; // [0] ready counter reset, [1] upload complete, [2] run words final
    DevBuf<unsigned long long> reads_ready; // bytes of the read slab that have arrived
    unsigned long long* h_ready = nullptr; // page-locked: the values copied into reads_ready after each chunk
    bool upload_in_flight = false; // seed_kernel has to wait on reads_ready
    uint32_t* early_runs = nullptr; // host destination of the run words, copied as soon as nwasm_kernel is done
    int64_t early_runs_cap = 0;
    bool early_runs_done = false;
    DevBuf<int> off_check; // offsets_check_kernel's result
    bool reported_only = false; // ma_b200_set_reported_only
    bool compacted = false; // the last run filled info_out / alns_out
    int64_t n_reported = 0;
    DevBuf<unsigned long long> soc_scratch; // socbuild_kernel -> harmonize_kernel
    DevBuf<int> soc_nmax;
    DevBuf<ReadInfo> info_out;
    DevBuf<DAln> alns_out;
};

static KswScore make_score( const ma_b200_params& p )
{
    KswScore s;
    s.match = p.match;
    s.mismatch = -p.mismatch;
    int q = p.gap, e = p.extend, q2 = p.gap2, e2 = p.extend2;
    s.qe_row0 = q + e;
    if( q2 + e2 < q + e ) // kswcpp_core.h:367-375
        std::swap( q, q2 ), std::swap( e, e2 );
    s.q = q, s.e = e, s.q2 = q2, s.e2 = e2;
    long long lt = e != e2 ? ( q2 - q ) / ( e - e2 ) - 1 : 0; // kswcpp_core.h:414-417
    if( q2 + e2 + lt * e2 > q + e + lt * e )
        ++lt;
    s.long_thres = (int)lt;
    s.long_diff = (int)( lt * ( e - e2 ) - ( q2 - q ) - e2 );
    s.min16 = std::min( { -p.mismatch, -p.gap, -p.extend, -p.gap2, -p.extend2 } ); // kswcpp.h:81
    int min_sc = std::min( -p.mismatch, 0 );
    s.early_return = ( -min_sc > 2 * ( q + e ) ) ? 1 : 0; // kswcpp_core.h:411-412
    return s;
}

#define MA_API_BEGIN                                                                                                   \
    if( !ctx )                                                                                                         \
        return MA_B200_EINVAL;                                                                                         \
    try                                                                                                                \
    {                                                                                                                  \
        MA_CUDA( cudaSetDevice( ctx->device ) );
#define MA_API_END                                                                                                     \
    }                                                                                                                  \
    catch( const ma::CudaError& e )                                                                                    \
    {                                                                                                                  \
        ctx->err = e.msg;                                                                                              \
        return MA_B200_ECUDA;                                                                                          \
    }                                                                                                                  \
    catch( const std::exception& e )                                                                                   \
    {                                                                                                                  \
        ctx->err = e.what( );                                                                                          \
        return MA_B200_EINVAL;                                                                                         \
    }                                                                                                                  \
    return MA_B200_OK;

extern "C" int ma_b200_params_preset( const char* name, ma_b200_params* p )
{
    if( !name || !p )
        return MA_B200_EINVAL;
    std::string s( name );
    for( auto& c : s )
        c = (char)tolower( c );
    s.erase( std::remove_if( s.begin( ), s.end( ), []( char c ) { return c == '_' || c == ' ' || c == '-'; } ),
             s.end( ) );
    memset( p, 0, sizeof( *p ) );
    // "Default" Presetting, parameter.h:621-880, and pGlobalParams, parameter.h:1032-1046
    p->match = 2, p->mismatch = 4, p->gap = 4, p->extend = 2, p->gap2 = 24, p->extend2 = 1, p->sv_penalty = 100;
    p->seeding_technique = 0, p->min_seed_length = 16, p->min_ambiguity = 0, p->max_ambiguity = 100;
    p->seed_drop_min_size = 15, p->seed_drop_factor = 0.005;
    p->max_num_soc = 30, p->min_num_soc = 1, p->soc_width = 0, p->rectangular_soc = 1;
    p->soc_score_drop = 0.1, p->harm_score_min = 18, p->harm_score_min_rel = 0.002;
    p->score_diff_tolerance = 0.0001, p->max_score_lookahead = 3, p->switch_qlen = 800;
    p->max_delta_dist = 0.1, p->min_delta_dist = 16;
    p->optimistic_gap_estimation = 1, p->gap_cost_cutting = 1, p->max_gap_area = 20;
    p->genome_size_disable = 10000000, p->disable_heuristics = 0;
    p->padding = 1000, p->bandwidth_ext = 512, p->min_bandwidth_gap = 20, p->zdrop = 200;
    p->srand_base = 0;
    p->report_n = 0, p->min_alignment_score = 75, p->max_supplementary_per_prim = 1, p->use_paired_reads = 0;
    p->max_overlap_supplementary = 0.1, p->paired_mean = 400, p->paired_std = 150, p->paired_bonus = 1.25;
    // ParameterSetManager(), parameter.h:1079-1104
    if( s == "default" )
        return MA_B200_OK;
    if( s == "illumina" || s == "illuminapaired" )
    {
        p->seeding_technique = 1, p->max_ambiguity = 500, p->min_num_soc = 10, p->max_num_soc = 20;
        p->use_paired_reads = s == "illuminapaired";
        return MA_B200_OK;
    }
    if( s == "pacbio" )
    {
        p->min_num_soc = 5, p->max_supplementary_per_prim = 100;
        return MA_B200_OK;
    }
    if( s == "nanopore" )
    {
        p->seeding_technique = 1, p->min_num_soc = 5, p->max_supplementary_per_prim = 100;
        return MA_B200_OK;
    }
    return MA_B200_EINVAL;
}

extern "C" int ma_b200_create( int device, ma_b200_ctx** out )
{
    if( !out )
        return MA_B200_EINVAL;
    *out = nullptr;
    int n = 0;
    if( cudaGetDeviceCount( &n ) != cudaSuccess || n <= 0 || device < 0 || device >= n )
        return MA_B200_ECUDA; // no CPU fallback: without a CUDA device there is no context
    ma_b200_ctx* ctx = new ma_b200_ctx( );
    ctx->device = device;
    try
    {
        MA_CUDA( cudaSetDevice( device ) );
        cudaDeviceProp prop;
        MA_CUDA( cudaGetDeviceProperties( &prop, device ) );
        ctx->num_sms = prop.multiProcessorCount;
        MA_CUDA( cudaStreamCreateWithFlags( &ctx->stream, cudaStreamNonBlocking ) );
        ctx->timer.init( );
        ma_b200_params_preset( "default", &ctx->params );
    }
    catch( const ma::CudaError& e )
    {
        fprintf( stderr, "ma_b200_create: %s\n", e.msg.c_str( ) );
        delete ctx;
        return MA_B200_ECUDA;
    }
    *out = ctx;
    return MA_B200_OK;
}

extern "C" int ma_b200_create_sibling( ma_b200_ctx* ctx, ma_b200_ctx** out )
{
    if( !ctx || !out )
        return MA_B200_EINVAL;
    if( !ctx->have_index )
    {
        ctx->err = "create_sibling: no index uploaded";
        return MA_B200_ESTATE;
    }
    const int rc = ma_b200_create( ctx->device, out );
    if( rc )
        return rc;
    ( *out )->params = ctx->params, ( *out )->reported_only = ctx->reported_only;
    ( *out )->index = ctx->index, ( *out )->have_index = true; // a view: the index slabs stay owned by ctx
    return MA_B200_OK;
}

extern "C" void ma_b200_destroy( ma_b200_ctx* ctx )
{
    if( !ctx )
        return;
    cudaSetDevice( ctx->device );
    if( ctx->shadow )
        ma_b200_destroy( ctx->shadow );
    if( ctx->stream )
        cudaStreamDestroy( ctx->stream );
    for( auto& e : ctx->binEv )
        for( auto& f : e )
            if( f )
                cudaEventDestroy( f );
    for( auto& e : ctx->ev )
        if( e )
            cudaEventDestroy( e );
    for( auto& e : ctx->ev_copy )
        if( e )
            cudaEventDestroy( e );
    if( ctx->copy_stream )
        cudaStreamDestroy( ctx->copy_stream );
    if( ctx->h_ready )
        cudaFreeHost( ctx->h_ready );
    delete ctx;
}

extern "C" const char* ma_b200_last_error( const ma_b200_ctx* ctx )
{
    return ctx ? ctx->err.c_str( ) : "null context";
}

extern "C" int64_t ma_b200_launch_count( const ma_b200_ctx* ctx )
{
    return ctx ? ctx->launches : 0;
}

extern "C" void* ma_b200_host_alloc( int64_t bytes )
{
    void* p = nullptr;
    if( bytes <= 0 || cudaHostAlloc( &p, (size_t)bytes, cudaHostAllocPortable ) != cudaSuccess )
    {
        cudaGetLastError( );
        return nullptr;
    }
    return p;
}

extern "C" void ma_b200_host_free( void* p )
{
    if( p )
        cudaFreeHost( p );
}

extern "C" int ma_b200_set_params( ma_b200_ctx* ctx, const ma_b200_params* params )
{
    if( !ctx || !params )
        return MA_B200_EINVAL;
    // values this implementation cannot honour are refused here instead of producing results that are neither of the
    // reference's modes (the reference's own parameter classes check ranges the same way, parameter.h:160-215)
    const ma_b200_params& p = *params;
    const char* why = nullptr;
    if( !p.rectangular_soc )
        why = "rectangular_soc = 0: the strand-split non-rectangular SoC of the SV presets (stripOfConsideration.h:97-112) "
              "is not implemented";
    else if( p.seeding_technique != 0 && p.seeding_technique != 1 )
        why = "seeding_technique must be 0 (maxSpan) or 1 (SMEMs)";
    else if( p.match <= 0 || p.mismatch < 0 || p.gap < 0 || p.extend < 0 || p.gap2 < 0 || p.extend2 < 0 )
        why = "scores: match must be positive, penalties non-negative";
    else if( p.max_num_soc < 1 || p.max_num_soc > MA_MAX_SETS_PER_READ )
        why = "max_num_soc must be in [1, 128]";
    else if( p.min_num_soc < 0 || p.max_ambiguity < 0 || p.min_ambiguity < 0 || p.min_seed_length < 0 )
        why = "min_num_soc, min/max_ambiguity and min_seed_length must be non-negative";
    else if( p.padding < 0 || p.bandwidth_ext < 1 || p.min_bandwidth_gap < 1 || p.max_gap_area < 0 )
        why = "padding, max_gap_area must be non-negative and the band widths positive";
    else if( p.report_n < 0 || p.max_supplementary_per_prim < 0 )
        why = "report_n and max_supplementary_per_prim must be non-negative";
    else if( p.use_paired_reads && !( p.paired_std > 0 ) )
        why = "paired_std must be positive with use_paired_reads";
    if( why )
    {
        ctx->err = std::string( "set_params: " ) + why;
        return MA_B200_EINVAL;
    }
    ctx->params = *params;
    return MA_B200_OK;
}

// ------------------------------------------------------------------------------------------------ DP batch
static const int kKswWindows[] = { 128, 256, 512, 1024, 2048 };

template <int W> static long long ksw_bin_grid( ma_b200_ctx* ctx, const KswHostBin& bin, long long tbBudget )
{
    const int warpsPerCta = KswClass<W>::WARPS;
    const size_t smem = KswSmemBytes<W>::value * warpsPerCta;
    MA_CUDA( cudaFuncSetAttribute( ksw_batch_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
    int perSm = 0;
    MA_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, ksw_batch_kernel<W>, 32 * warpsPerCta, smem ) );
    if( perSm < 1 )
        perSm = 1;
    long long grid = (long long)perSm * ctx->num_sms;
    grid = std::min<long long>( grid, ( (long long)bin.order.size( ) + warpsPerCta - 1 ) / warpsPerCta );
    const long long perCta = ( bin.tb_stride + 4ll * bin.cig_stride ) * warpsPerCta;
    if( perCta > 0 )
        grid = std::min<long long>( grid, std::max<long long>( 1, tbBudget / perCta ) );
    return std::max<long long>( grid, 1 );
}

template <int W> static void launch_ksw_bin( ma_b200_ctx* ctx, const KswBatchArgs& A, long long grid )
{
    const size_t smem = KswSmemBytes<W>::value * KswClass<W>::WARPS;
    ksw_batch_kernel<W><<<(unsigned)grid, 32 * KswClass<W>::WARPS, smem, ctx->stream>>>( A );
    MA_CUDA( cudaGetLastError( ) );
    ctx->launches++;
}

// ksw_qs_kernel: qs = 1 + (blocks - 1) * 2 + right-aligned
template <int NB, bool LEFT>
static long long ksw_qs_grid_t( ma_b200_ctx* ctx, long long nTasks, long long tbStride, int cigStride, long long tbBudget )
{
    int perSm = 0;
    MA_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, ksw_qs_kernel<NB, LEFT>, 32 * MA_QS_WARPS, 0 ) );
    if( perSm < 1 )
        perSm = 1;
    long long grid = (long long)perSm * ctx->num_sms;
    grid = std::min<long long>( grid, ( nTasks + MA_QS_WARPS - 1 ) / MA_QS_WARPS );
    const long long perCta = ( tbStride + 4ll * cigStride ) * MA_QS_WARPS;
    if( perCta > 0 )
        grid = std::min<long long>( grid, std::max<long long>( 1, tbBudget / perCta ) );
    return std::max<long long>( grid, 1 );
}
static long long ksw_qs_grid( ma_b200_ctx* ctx, int qs, long long nTasks, long long tbStride, int cigStride,
                              long long tbBudget )
{
    switch( qs )
    {
        case 1: return ksw_qs_grid_t<1, true>( ctx, nTasks, tbStride, cigStride, tbBudget );
        case 2: return ksw_qs_grid_t<1, false>( ctx, nTasks, tbStride, cigStride, tbBudget );
        case 3: return ksw_qs_grid_t<2, true>( ctx, nTasks, tbStride, cigStride, tbBudget );
        case 4: return ksw_qs_grid_t<2, false>( ctx, nTasks, tbStride, cigStride, tbBudget );
        case 5: return ksw_qs_grid_t<3, true>( ctx, nTasks, tbStride, cigStride, tbBudget );
        default: return ksw_qs_grid_t<3, false>( ctx, nTasks, tbStride, cigStride, tbBudget );
    }
}
static void launch_ksw_qs( ma_b200_ctx* ctx, int qs, const KswBatchArgs& A, long long grid )
{
    const unsigned g = (unsigned)grid, b = 32 * MA_QS_WARPS;
    switch( qs )
    {
        case 1: ksw_qs_kernel<1, true><<<g, b, 0, ctx->stream>>>( A ); break;
        case 2: ksw_qs_kernel<1, false><<<g, b, 0, ctx->stream>>>( A ); break;
        case 3: ksw_qs_kernel<2, true><<<g, b, 0, ctx->stream>>>( A ); break;
        case 4: ksw_qs_kernel<2, false><<<g, b, 0, ctx->stream>>>( A ); break;
        case 5: ksw_qs_kernel<3, true><<<g, b, 0, ctx->stream>>>( A ); break;
        default: ksw_qs_kernel<3, false><<<g, b, 0, ctx->stream>>>( A ); break;
    }
    MA_CUDA( cudaGetLastError( ) );
    ctx->launches++;
}
static bool use_tiny( )
{
    static const bool b = !( getenv( "MA_B200_NO_TINY" ) && atoi( getenv( "MA_B200_NO_TINY" ) ) != 0 );
    return b;
}
static int no_bx_bits( )
{ // A/B measurements: MA_B200_NO_BX=1 runs the scalar exact mode instead of the packed banded one (ksw_bx.cuh),
  // MA_B200_NO_BX=2 skips the half2 two-row mode (ksw_rows_p2x2), 4 the register-resident narrow-band mode (ksw_bn.cuh); bits add
    static const int b = getenv( "MA_B200_NO_BX" ) ? atoi( getenv( "MA_B200_NO_BX" ) ) : 0;
    return b;
}
static bool use_qs( )
{
    static const bool b = !( getenv( "MA_B200_NO_QS" ) && atoi( getenv( "MA_B200_NO_QS" ) ) != 0 );
    return b;
}

// bins the (device-tagged) tasks: ksw_qs_kernel where it applies, else the window classes of ksw_batch_kernel.
// A band wider than the largest window fails that TASK (status 2), not the batch.
static void ksw_plan( ma_b200_ctx* ctx, const std::vector<KswTask>& tasks, std::vector<int>& tooWide )
{
    const KswScore score = make_score( ctx->params );
    const int64_t n = (int64_t)tasks.size( );
    ctx->ksw_bins.clear( );
    for( int W : kKswWindows )
        ctx->ksw_bins.push_back( KswHostBin{ W, { }, 0, 0, 0 } );
    for( int qs = 1; qs <= 6; qs++ )
        ctx->ksw_bins.push_back( KswHostBin{ 0, { }, 0, 0, qs } );
    long long bound = 0;
    std::vector<long long> cost( n );
    for( int64_t i = 0; i < n; i++ )
    {
        const auto& t = tasks[ i ];
        if( t.qlen < 0 || t.tlen < 0 )
            throw std::runtime_error( "ksw task with negative length" );
        const int nc = ksw_ncol16( t.qlen, t.tlen, t.w );
        const long long rows = (long long)t.qlen + t.tlen;
        int b = -1;
        long long tbBytes = ( rows * nc + 255 ) & ~255ll;
        const int nb = use_qs( ) ? ksw_qs_class( score, t.qlen, t.tlen, t.w, t.tag ) : 0;
        if( nb > 0 )
        {
            b = (int)( sizeof( kKswWindows ) / sizeof( int ) ) + ( nb - 1 ) * 2 + ( ( t.flag & MA_KSW_RIGHT ) ? 1 : 0 );
            tbBytes = ksw_qs_tb_bytes( nb, t.qlen, t.tlen, t.w );
        }
        else
            for( size_t k = 0; k < sizeof( kKswWindows ) / sizeof( int ); k++ )
                if( ksw_class_cap( ctx->ksw_bins[ k ].W ) >= nc + 48 )
                {
                    b = (int)k;
                    break;
                }
        if( b < 0 )
        {
            tooWide.push_back( (int)i );
            continue;
        }
        auto& bin = ctx->ksw_bins[ b ];
        bin.order.push_back( (int)i );
        bin.tb_stride = std::max( bin.tb_stride, tbBytes );
        bin.cig_stride = std::max<int>( bin.cig_stride, (int)( ( rows + 2 + 63 ) & ~63ll ) );
        cost[ i ] = rows * nc;
        bound += rows + 2;
    }
    for( auto& bin : ctx->ksw_bins ) // longest first: the dynamic queue then load-balances the tail
        std::stable_sort( bin.order.begin( ), bin.order.end( ),
                          [ & ]( int a, int b ) { return cost[ a ] > cost[ b ]; } );
    ctx->ksw_cigar_bound = bound;
}

extern "C" int ma_b200_ksw_set_extension_only( ma_b200_ctx* ctx, int32_t on )
{
    if( !ctx )
        return MA_B200_EINVAL;
    ctx->ksw_extension_only = on != 0;
    return MA_B200_OK;
}

extern "C" int ma_b200_ksw_upload( ma_b200_ctx* ctx, int64_t n, const ma_b200_ksw_task* tasks, const uint8_t* seq,
                                   int64_t seq_bytes )
{
    MA_API_BEGIN
    if( n < 0 || ( n > 0 && ( !tasks || !seq ) ) || n > 0x7fffffff )
        throw std::runtime_error( "ksw_upload: bad arguments" );
    for( int64_t i = 0; i < n; i++ )
        if( tasks[ i ].qoff < 0 || tasks[ i ].toff < 0 || tasks[ i ].qoff + tasks[ i ].qlen > seq_bytes ||
            tasks[ i ].toff + tasks[ i ].tlen > seq_bytes )
            throw std::runtime_error( "ksw_upload: task sequence range outside the slab" );
    static_assert( sizeof( ma_b200_ksw_task ) == sizeof( KswTask ), "task layout" );
    static_assert( sizeof( ma_b200_ksw_result ) == sizeof( KswOut ), "result layout" );
    // `tag` is the caller's cookie; on the device the field carries the internal addressing mode (0 = byte slab)
    std::vector<KswTask>& vTasks = ctx->ksw_host_tasks;
    vTasks.resize( (size_t)n );
    if( n > 0 )
        memcpy( vTasks.data( ), tasks, n * sizeof( KswTask ) );
    for( auto& t : vTasks )
        t.tag = ( ctx->ksw_extension_only && ( t.flag & MA_KSW_EXTZ_ONLY ) ) ? MA_TASK_EARLYSTOP : 0;
    std::vector<int> tooWide;
    ksw_plan( ctx, vTasks, tooWide );
    ctx->ksw_n = n;
    ctx->ksw_tasks.reserve( (size_t)n + 1 );
    ctx->ksw_seq.reserve( (size_t)seq_bytes + 1 );
    ctx->ksw_out.reserve( (size_t)n + 1 );
    ctx->ksw_order.reserve( (size_t)n + 1 );
    ctx->ksw_redo.reserve( (size_t)n + 2 );
    ctx->ksw_ctrl.reserve( 128 );
    if( n > 0 )
    {
        MA_CUDA( cudaMemcpyAsync( ctx->ksw_tasks.p, vTasks.data( ), n * sizeof( KswTask ), cudaMemcpyHostToDevice,
                                  ctx->stream ) );
        MA_CUDA( cudaMemcpyAsync( ctx->ksw_seq.p, seq, seq_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
        size_t o = 0;
        for( auto& bin : ctx->ksw_bins )
        {
            if( !bin.order.empty( ) )
                MA_CUDA( cudaMemcpyAsync( ctx->ksw_order.p + o, bin.order.data( ), bin.order.size( ) * sizeof( int ),
                                          cudaMemcpyHostToDevice, ctx->stream ) );
            o += bin.order.size( );
        }
        if( !tooWide.empty( ) )
        { // status 2: band wider than the largest supported window (2000 columns); the rest of the batch is computed
            KswOut bad;
            memset( &bad, 0, sizeof( bad ) );
            bad.max_q = bad.max_t = bad.mqe_t = bad.mte_q = -1, bad.status = 2;
            for( int i : tooWide )
                MA_CUDA( cudaMemcpyAsync( ctx->ksw_out.p + i, &bad, sizeof( bad ), cudaMemcpyHostToDevice, ctx->stream ) );
            MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
        }
    }
    // cigar slab: start with a typical size; ksw_run re-runs with the exact bound if it overflows
    ctx->ksw_cigar_cap = std::min<long long>( ctx->ksw_cigar_bound, std::max<long long>( 48 * n, 1 << 16 ) );
    ctx->ksw_cigar.reserve( (size_t)ctx->ksw_cigar_cap + 1 );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

static long long ksw_host_grid( ma_b200_ctx* ctx, const KswHostBin& bin, long long budget )
{
    if( bin.order.empty( ) )
        return 0;
    if( bin.qs )
        return ksw_qs_grid( ctx, bin.qs, (long long)bin.order.size( ), bin.tb_stride, bin.cig_stride, budget );
    switch( bin.W )
    {
        case 128: return ksw_bin_grid<128>( ctx, bin, budget );
        case 256: return ksw_bin_grid<256>( ctx, bin, budget );
        case 512: return ksw_bin_grid<512>( ctx, bin, budget );
        case 1024: return ksw_bin_grid<1024>( ctx, bin, budget );
        default: return ksw_bin_grid<2048>( ctx, bin, budget );
    }
}
static void ksw_host_launch( ma_b200_ctx* ctx, const KswHostBin& bin, KswBatchArgs& A, long long grid )
{
    A.n = (int)bin.order.size( );
    A.tb_stride = bin.tb_stride, A.cigscratch_stride = bin.cig_stride;
    MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p + 16, 0, sizeof( unsigned long long ), ctx->stream ) );
    if( bin.qs )
        return launch_ksw_qs( ctx, bin.qs, A, grid );
    switch( bin.W )
    {
        case 128: launch_ksw_bin<128>( ctx, A, grid ); break;
        case 256: launch_ksw_bin<256>( ctx, A, grid ); break;
        case 512: launch_ksw_bin<512>( ctx, A, grid ); break;
        case 1024: launch_ksw_bin<1024>( ctx, A, grid ); break;
        default: launch_ksw_bin<2048>( ctx, A, grid ); break;
    }
}

static int ksw_run_once( ma_b200_ctx* ctx )
{
    MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p, 0, 128 * sizeof( unsigned long long ), ctx->stream ) );
    MA_CUDA( cudaMemsetAsync( ctx->ksw_redo.p, 0, sizeof( int ), ctx->stream ) );
    const KswScore score = make_score( ctx->params );
    const long long budget = 48ll << 30; // traceback + cigar scratch of all resident warps (180 GB of HBM per GPU)
    KswBatchArgs A;
    A.tasks = ctx->ksw_tasks.p;
    A.seq = ctx->ksw_seq.p;
    A.pac = nullptr, A.fwd_len = 0;
    A.out = ctx->ksw_out.p;
    A.cigar = ctx->ksw_cigar.p;
    A.cigar_cap = ctx->ksw_cigar_cap;
    A.cigar_cursor = ctx->ksw_ctrl.p;
    A.next = (int*)( ctx->ksw_ctrl.p + 16 );
    A.error = (int*)( ctx->ksw_ctrl.p + 32 );
    A.cells_total = nullptr;
    A.score = score;
    A.qsk[ 0 ] = ksw_qs_make_k( score, true ), A.qsk[ 1 ] = ksw_qs_make_k( score, false );
    A.bxk[ 0 ] = ksw_bx_make_k( score, true ), A.bxk[ 1 ] = ksw_bx_make_k( score, false );
    A.no_bx = no_bx_bits( );
    A.redo_count = nullptr, A.redo_tb = nullptr, A.redo_cig = nullptr;
    A.redo_n = ctx->ksw_redo.p, A.redo_order = ctx->ksw_redo.p + 1, A.redo_cap = ctx->ksw_n;
    // two phases: ksw_qs_kernel first (it may hand problems over), then ksw_batch_kernel. The per-warp scratch is sized
    // per phase (DevBuf::reserve may free + reallocate: only between phases, after a synchronisation).
    std::vector<KswHostBin> redoBins;
    for( int phase = 0; phase < 2; phase++ )
    {
        std::vector<const KswHostBin*> bins;
        std::vector<const int*> orders;
        size_t o = 0;
        for( auto& bin : ctx->ksw_bins )
        {
            if( !bin.order.empty( ) && ( bin.qs != 0 ) == ( phase == 0 ) )
                bins.push_back( &bin ), orders.push_back( ctx->ksw_order.p + o );
            o += bin.order.size( );
        }
        if( phase == 1 )
        { // problems handed over by ksw_qs_kernel, binned by window class
            int nRedo = 0;
            MA_CUDA( cudaMemcpyAsync( &nRedo, ctx->ksw_redo.p, sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
            MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
            if( nRedo > 0 )
            {
                std::vector<int> ids( (size_t)nRedo );
                MA_CUDA( cudaMemcpyAsync( ids.data( ), ctx->ksw_redo.p + 1, nRedo * sizeof( int ), cudaMemcpyDeviceToHost,
                                          ctx->stream ) );
                MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
                std::sort( ids.begin( ), ids.end( ) );
                for( int W : kKswWindows )
                    redoBins.push_back( KswHostBin{ W, { }, 0, 0, 0 } );
                for( int i : ids )
                {
                    const KswTask& t = ctx->ksw_host_tasks[ i ];
                    const int nc = ksw_ncol16( t.qlen, t.tlen, t.w );
                    const long long rows = (long long)t.qlen + t.tlen;
                    for( auto& rb : redoBins )
                        if( ksw_class_cap( rb.W ) >= nc + 48 )
                        {
                            rb.order.push_back( i );
                            rb.tb_stride = std::max( rb.tb_stride, ( rows * nc + 255 ) & ~255ll );
                            rb.cig_stride = std::max<int>( rb.cig_stride, (int)( ( rows + 2 + 63 ) & ~63ll ) );
                            break;
                        }
                }
                size_t ro = 1;
                for( auto& rb : redoBins )
                {
                    if( rb.order.empty( ) )
                        continue;
                    MA_CUDA( cudaMemcpyAsync( ctx->ksw_redo.p + ro, rb.order.data( ), rb.order.size( ) * sizeof( int ),
                                              cudaMemcpyHostToDevice, ctx->stream ) );
                    bins.push_back( &rb ), orders.push_back( ctx->ksw_redo.p + ro );
                    ro += rb.order.size( );
                }
                MA_CUDA( cudaStreamSynchronize( ctx->stream ) ); // the host vectors stay alive, but keep it simple
            }
        }
        std::vector<long long> grids;
        size_t tbNeed = 0, csNeed = 0;
        for( const KswHostBin* bin : bins )
        {
            const long long g = ksw_host_grid( ctx, *bin, budget );
            grids.push_back( g );
            const int wpc = bin->qs ? MA_QS_WARPS : ( bin->W == 1024 ? KswClass<1024>::WARPS : MA_KSW_WARPS );
            tbNeed = std::max<size_t>( tbNeed, (size_t)( g * wpc * bin->tb_stride ) );
            csNeed = std::max<size_t>( csNeed, (size_t)( g * wpc * bin->cig_stride ) );
        }
        ctx->ksw_tb.reserve( tbNeed + 256 );
        ctx->ksw_cigscratch.reserve( csNeed + 64 );
        A.tb = ctx->ksw_tb.p, A.cigscratch = ctx->ksw_cigscratch.p;
        for( size_t k = 0; k < bins.size( ); k++ )
        {
            A.order = orders[ k ];
            ksw_host_launch( ctx, *bins[ k ], A, grids[ k ] );
        }
    }
    unsigned long long ctrl[ 64 ];
    MA_CUDA( cudaMemcpyAsync( ctrl, ctx->ksw_ctrl.p, sizeof( ctrl ), cudaMemcpyDeviceToHost, ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->ksw_cigar_used = ctrl[ 0 ];
    return (int)ctrl[ 32 ];
}

extern "C" int ma_b200_ksw_run( ma_b200_ctx* ctx, float* kernel_ms )
{
    NvtxRange nvtxCall( "ma_b200_ksw_run" );
    MA_API_BEGIN
    if( kernel_ms )
        *kernel_ms = 0;
    if( ctx->ksw_n == 0 )
    {
        ctx->ksw_cigar_used = 0;
        return MA_B200_OK;
    }
    ctx->timer.start( ctx->stream );
    int err = ksw_run_once( ctx );
    if( err )
    { // cigar slab overflow: grow to the exact upper bound and run again
        ctx->ksw_cigar_cap = ctx->ksw_cigar_bound;
        ctx->ksw_cigar.reserve( (size_t)ctx->ksw_cigar_cap + 1 );
        err = ksw_run_once( ctx );
        if( err )
            throw std::runtime_error( "ksw_run: cigar slab overflow at the exact bound (internal error)" );
    }
    const float ms = ctx->timer.stop( ctx->stream );
    if( kernel_ms )
        *kernel_ms = ms;
    MA_API_END
}

extern "C" int ma_b200_ksw_download( ma_b200_ctx* ctx, ma_b200_ksw_result* results, uint32_t* cigar,
                                     int64_t cigar_cap_words, int64_t* cigar_words )
{
    MA_API_BEGIN
    if( cigar_words )
        *cigar_words = (int64_t)ctx->ksw_cigar_used;
    if( ctx->ksw_n == 0 )
        return MA_B200_OK;
    if( !results )
        throw std::runtime_error( "ksw_download: null results" );
    if( (int64_t)ctx->ksw_cigar_used > cigar_cap_words )
    {
        ctx->err = "ksw_download: cigar slab too small";
        return MA_B200_ENOMEM;
    }
    MA_CUDA( cudaMemcpyAsync( results, ctx->ksw_out.p, ctx->ksw_n * sizeof( KswOut ), cudaMemcpyDeviceToHost,
                              ctx->stream ) );
    if( ctx->ksw_cigar_used > 0 )
        MA_CUDA( cudaMemcpyAsync( cigar, ctx->ksw_cigar.p, ctx->ksw_cigar_used * sizeof( unsigned int ),
                                  cudaMemcpyDeviceToHost, ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

extern "C" int ma_b200_ksw_batch( ma_b200_ctx* ctx, int64_t n, const ma_b200_ksw_task* tasks, const uint8_t* seq,
                                  int64_t seq_bytes, ma_b200_ksw_result* results, uint32_t* cigar,
                                  int64_t cigar_cap_words, int64_t* cigar_words )
{
    int rc = ma_b200_ksw_upload( ctx, n, tasks, seq, seq_bytes );
    if( rc )
        return rc;
    rc = ma_b200_ksw_run( ctx, nullptr );
    if( rc )
        return rc;
    return ma_b200_ksw_download( ctx, results, cigar, cigar_cap_words, cigar_words );
}

// ------------------------------------------------------------------------------------------------ index
// reference occ blocks -> bit-plane blocks, one thread per 128-symbol block (the trailing counter block, which sits
// right behind a short last data block, is never taken for symbols)
__global__ void index_relayout_kernel( const unsigned int* ref, long long n_words, long long N, long long nblk,
                                       unsigned int* planes )
{
    for( long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk;
         b += (long long)gridDim.x * blockDim.x )
    {
        unsigned int in[ 16 ], out[ 16 ];
        for( int j = 0; j < 16; j++ )
        {
            const long long idx = 16 * b + j;
            const bool ok = idx < n_words && ( j < 8 || 128 * b + 16 * ( j - 8 ) < N );
            in[ j ] = ok ? ref[ idx ] : 0u;
        }
        relayout_block( in, out );
        for( int j = 0; j < 16; j++ )
            planes[ 16 * b + j ] = out[ j ];
    }
}

static void index_make_planes( ma_b200_ctx* ctx, long long n_words, long long ref_len )
{
    const long long nblk = ( ref_len + 127 ) / 128;
    ctx->ix_bwtp.reserve( (size_t)( nblk + 2 ) * 4 );
    MA_CUDA( cudaMemsetAsync( ctx->ix_bwtp.p + nblk * 4, 0, 2 * 64, ctx->stream ) );
    const int grid = (int)std::min<long long>( ( nblk + 255 ) / 256, (long long)ctx->num_sms * 8 );
    index_relayout_kernel<<<grid, 256, 0, ctx->stream>>>( (const unsigned int*)ctx->ix_bwt.p, n_words, ref_len, nblk,
                                                          (unsigned int*)ctx->ix_bwtp.p );
    MA_CUDA( cudaGetLastError( ) );
    ctx->launches++;
    // Optional (MA_B200_L2_WINDOW=<hit ratio in percent>): an L2 access-policy window over the occurrence table, the
    // north_star's "top BWT levels pinned in L2". The blocks of the shallow extension depths are NOT contiguous (the 4^d
    // intervals of depth d lie all over the table), so the window can only cover the table as a whole, with a hit
    // ratio that keeps the persisting part within the L2 set-aside. Measured in DESIGN.md §4.2.
    if( const char* e = getenv( "MA_B200_L2_WINDOW" ) )
    {
        const int pct = atoi( e );
        if( pct > 0 )
        {
            cudaDeviceProp prop;
            MA_CUDA( cudaGetDeviceProperties( &prop, ctx->device ) );
            const size_t bytes = (size_t)nblk * 64;
            const size_t setAside = std::min<size_t>( (size_t)prop.persistingL2CacheMaxSize, bytes );
            MA_CUDA( cudaDeviceSetLimit( cudaLimitPersistingL2CacheSize, setAside ) );
            cudaStreamAttrValue attr;
            memset( &attr, 0, sizeof( attr ) );
            attr.accessPolicyWindow.base_ptr = ctx->ix_bwtp.p;
            attr.accessPolicyWindow.num_bytes = std::min<size_t>( bytes, (size_t)prop.accessPolicyMaxWindowSize );
            attr.accessPolicyWindow.hitRatio = std::min( 1.0f, pct / 100.0f );
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            MA_CUDA( cudaStreamSetAttribute( ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr ) );
            fprintf( stderr, "ma_b200: L2 window over %zu MB of the occurrence table (max window %d MB), set-aside %zu MB, hit ratio %.2f\n",
                     bytes >> 20, prop.accessPolicyMaxWindowSize >> 20, setAside >> 20, attr.accessPolicyWindow.hitRatio );
        }
    }
}

extern "C" int ma_b200_index_upload( ma_b200_ctx* ctx, const uint32_t* bwt_words, int64_t n_words, const int64_t* L2,
                                     int64_t primary, int64_t ref_len, const int64_t* sa, int64_t n_sa,
                                     int32_t sa_intv, const uint8_t* pac, int64_t n_pac_bytes, int64_t fwd_len,
                                     const int64_t* contig_start, const int64_t* contig_len, int32_t n_contigs )
{
    NvtxRange nvtxCall( "ma_b200_index_upload" );
    MA_API_BEGIN
    if( !bwt_words || !L2 || !sa || !pac || !contig_start || !contig_len || n_contigs <= 0 || n_words <= 0 ||
        sa_intv <= 0 || ( sa_intv & ( sa_intv - 1 ) ) || ref_len != 2 * fwd_len ||
        n_pac_bytes < ( fwd_len + 3 ) / 4 || n_sa < ( ref_len + sa_intv ) / sa_intv )
        throw std::runtime_error( "index_upload: inconsistent arguments" );
    ctx->ix_bwt.reserve( (size_t)n_words / 4 + 8 );
    ctx->ix_sa.reserve( (size_t)n_sa + 1 );
    ctx->ix_pac.reserve( (size_t)n_pac_bytes + 16 );
    ctx->ix_contigs.reserve( (size_t)2 * n_contigs );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_bwt.p, bwt_words, n_words * 4, cudaMemcpyHostToDevice, ctx->stream ) );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_sa.p, sa, n_sa * 8, cudaMemcpyHostToDevice, ctx->stream ) );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_pac.p, pac, n_pac_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_contigs.p, contig_start, n_contigs * 8, cudaMemcpyHostToDevice, ctx->stream ) );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_contigs.p + n_contigs, contig_len, n_contigs * 8, cudaMemcpyHostToDevice,
                              ctx->stream ) );
    index_make_planes( ctx, n_words, ref_len );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    DevIndex& I = ctx->index;
    I.bwt = ctx->ix_bwtp.p, I.sa = ctx->ix_sa.p, I.pac = ctx->ix_pac.p;
    I.contig_start = ctx->ix_contigs.p, I.contig_len = ctx->ix_contigs.p + n_contigs;
    for( int i = 0; i < 5; i++ )
        I.L2[ i ] = L2[ i ];
    I.primary = primary, I.ref_len = ref_len, I.fwd_len = fwd_len, I.sa_intv = sa_intv, I.n_contigs = n_contigs;
    ctx->ix_words = n_words, ctx->ix_nsa = n_sa, ctx->ix_npac = n_pac_bytes;
    ctx->have_index = true;
    MA_API_END
}

extern "C" int ma_b200_index_build( ma_b200_ctx* ctx, const uint8_t* fwd, int64_t fwd_len, const int64_t* contig_start,
                                    const int64_t* contig_len, int32_t n_contigs )
{
    NvtxRange nvtxCall( "ma_b200_index_build" );
    MA_API_BEGIN
    if( !fwd || fwd_len <= 0 || !contig_start || !contig_len || n_contigs <= 0 )
        throw std::runtime_error( "index_build: bad arguments" );
    ctx->have_index = false;
    // Small texts: whole suffix array by prefix doubling. From 2^31 - 1 suffixes on (or on request: MA_B200_IB_LARGE=1,
    // chunk size in suffixes via MA_B200_IB_CHUNK — the tests run the bucketed builder on small genomes that way) the
    // suffixes are sorted bucket by bucket (index_build.cuh build_index_gpu_large).
    const char* sLarge = getenv( "MA_B200_IB_LARGE" );
    const char* sChunk = getenv( "MA_B200_IB_CHUNK" );
    const bool bLarge = 2 * fwd_len >= 0x7fffffffll || ( sLarge && atoi( sLarge ) != 0 );
    IndexBuildResult R =
        bLarge ? build_index_gpu_large( ctx->stream, ctx->num_sms, fwd, fwd_len, sChunk ? atoll( sChunk ) : 0,
                                        ctx->ix_bwt, ctx->ix_sa, ctx->ix_pac, ctx->launches )
               : build_index_gpu( ctx->stream, ctx->num_sms, fwd, fwd_len, ctx->ix_bwt, ctx->ix_sa, ctx->ix_pac,
                                  ctx->launches );
    ctx->ix_contigs.reserve( (size_t)2 * n_contigs );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_contigs.p, contig_start, n_contigs * 8, cudaMemcpyHostToDevice, ctx->stream ) );
    MA_CUDA( cudaMemcpyAsync( ctx->ix_contigs.p + n_contigs, contig_len, n_contigs * 8, cudaMemcpyHostToDevice,
                              ctx->stream ) );
    index_make_planes( ctx, R.n_words, 2 * fwd_len );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    DevIndex& I = ctx->index;
    I.bwt = ctx->ix_bwtp.p, I.sa = ctx->ix_sa.p, I.pac = ctx->ix_pac.p;
    I.contig_start = ctx->ix_contigs.p, I.contig_len = ctx->ix_contigs.p + n_contigs;
    for( int i = 0; i < 5; i++ )
        I.L2[ i ] = R.L2[ i ];
    I.primary = R.primary, I.ref_len = 2 * fwd_len, I.fwd_len = fwd_len, I.sa_intv = 32, I.n_contigs = n_contigs;
    ctx->ix_words = R.n_words, ctx->ix_nsa = R.n_sa, ctx->ix_npac = R.n_pac;
    ctx->have_index = true;
    MA_API_END
}

extern "C" int ma_b200_index_sizes( ma_b200_ctx* ctx, int64_t* n_words, int64_t* n_sa, int64_t* n_pac_bytes,
                                    int64_t* primary, int64_t* L2 )
{
    if( !ctx || !ctx->have_index )
        return MA_B200_ESTATE;
    if( n_words )
        *n_words = ctx->ix_words;
    if( n_sa )
        *n_sa = ctx->ix_nsa;
    if( n_pac_bytes )
        *n_pac_bytes = ctx->ix_npac;
    if( primary )
        *primary = ctx->index.primary;
    if( L2 )
        for( int i = 0; i < 5; i++ )
            L2[ i ] = ctx->index.L2[ i ];
    return MA_B200_OK;
}

extern "C" int ma_b200_index_download( ma_b200_ctx* ctx, uint32_t* bwt_words, int64_t* sa, uint8_t* pac )
{
    MA_API_BEGIN
    if( !ctx->have_index )
    {
        ctx->err = "no index";
        return MA_B200_ESTATE;
    }
    if( bwt_words )
        MA_CUDA( cudaMemcpyAsync( bwt_words, ctx->ix_bwt.p, ctx->ix_words * 4, cudaMemcpyDeviceToHost, ctx->stream ) );
    if( sa )
        MA_CUDA( cudaMemcpyAsync( sa, ctx->ix_sa.p, ctx->ix_nsa * 8, cudaMemcpyDeviceToHost, ctx->stream ) );
    if( pac )
        MA_CUDA( cudaMemcpyAsync( pac, ctx->ix_pac.p, ctx->ix_npac, cudaMemcpyDeviceToHost, ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

// ------------------------------------------------------------------------------------------------ alignment path
static SeedParams make_seed_params( const ma_b200_params& p )
{
    return SeedParams{ p.seeding_technique, p.min_ambiguity, p.max_ambiguity, p.min_seed_length,
                       p.seed_drop_min_size, p.seed_drop_factor, p.disable_heuristics, p.genome_size_disable };
}
static HarmParams make_harm_params( const ma_b200_params& p )
{
    return HarmParams{ p.match, p.gap, p.extend, p.sv_penalty, p.max_num_soc, p.min_num_soc, p.soc_width,
                       p.rectangular_soc, p.soc_score_drop, p.harm_score_min, p.harm_score_min_rel,
                       p.score_diff_tolerance, p.max_score_lookahead, p.switch_qlen, p.max_delta_dist,
                       p.min_delta_dist, p.optimistic_gap_estimation, p.gap_cost_cutting, p.disable_heuristics,
                       p.genome_size_disable };
}
static NwParams make_nw_params( const ma_b200_params& p )
{
    return NwParams{ p.match, p.mismatch, p.gap, p.extend, p.sv_penalty, p.max_gap_area, p.padding,
                     p.bandwidth_ext, p.min_bandwidth_gap, p.zdrop, make_score( p ).early_return ? 0 : 1 };
}

static std::mutex& device_turnstile( int device )
{
    static std::mutex m[ 64 ];
    return m[ device & 63 ];
}

static void ensure_copy_stream( ma_b200_ctx* ctx )
{
    if( ctx->copy_stream )
        return;
    MA_CUDA( cudaStreamCreateWithFlags( &ctx->copy_stream, cudaStreamNonBlocking ) );
    for( auto& e : ctx->ev_copy )
        MA_CUDA( cudaEventCreateWithFlags( &e, cudaEventDisableTiming ) );
    MA_CUDA( cudaHostAlloc( (void**)&ctx->h_ready, 64 * sizeof( unsigned long long ), cudaHostAllocDefault ) );
}

// Upload of a batch. overlapped: the bases go up in chunks on the copy stream with a counter of arrived reads behind
// every chunk (seed_kernel waits on it), the offsets are validated on the host while the copies run, and the call
// returns without waiting for them (ma_b200_align_batch, one-shot form).
static void align_upload_impl( ma_b200_ctx* ctx, int64_t n_reads, const uint8_t* reads, const int64_t* offsets,
                               bool overlapped )
{
    if( n_reads < 0 || n_reads > 0x7ffffff0 || ( n_reads > 0 && ( !reads || !offsets ) ) )
        throw std::runtime_error( "align_upload: bad arguments" );
    ctx->upload_in_flight = false;
    const int64_t total = n_reads ? offsets[ n_reads ] : 0;
    bool bad = total < 0 || ( n_reads && offsets[ 0 ] < 0 );
    const int nChunks = 16;
    if( !overlapped || n_reads < 256 || total < ( 1 << 15 ) )
        overlapped = false;
    if( overlapped && !bad )
    {
        ensure_copy_stream( ctx );
        ctx->reads.reserve( (size_t)total + 256 );
        ctx->read_off.reserve( (size_t)n_reads + 1 );
        ctx->reads_ready.reserve( 1 );
        cudaStream_t cs = ctx->copy_stream;
        MA_CUDA( cudaMemsetAsync( ctx->reads_ready.p, 0, sizeof( unsigned long long ), cs ) );
        MA_CUDA( cudaMemcpyAsync( ctx->read_off.p, offsets, ( n_reads + 1 ) * 8, cudaMemcpyHostToDevice, cs ) );
        MA_CUDA( cudaEventRecord( ctx->ev_copy[ 0 ], cs ) );
        int64_t b0 = 0;
        for( int c = 0; c < nChunks; c++ )
        { // chunks end on 128-byte lines of the slab; the counter behind the last one covers the rounded-up end
            int64_t b1 = c + 1 < nChunks ? ( ( total / nChunks * ( c + 1 ) ) + 127 ) & ~(int64_t)127 : total;
            b1 = std::min( b1, total );
            if( b1 > b0 )
                MA_CUDA( cudaMemcpyAsync( ctx->reads.p + b0, reads + b0, b1 - b0, cudaMemcpyHostToDevice, cs ) );
            ctx->h_ready[ c ] = c + 1 < nChunks ? (unsigned long long)b1 : (unsigned long long)total + 128;
            MA_CUDA( cudaMemcpyAsync( ctx->reads_ready.p, ctx->h_ready + c, sizeof( unsigned long long ),
                                      cudaMemcpyHostToDevice, cs ) );
            b0 = b1;
        }
        MA_CUDA( cudaEventRecord( ctx->ev_copy[ 1 ], cs ) );
        MA_CUDA( cudaStreamWaitEvent( ctx->stream, ctx->ev_copy[ 0 ], 0 ) ); // kernels: after the reset + the offsets
        ctx->upload_in_flight = true;
    }
    int maxL = 0;
    if( ctx->upload_in_flight )
    { // the offsets are checked where they already are: one pass of a small kernel instead of a host loop over n reads
        ctx->off_check.reserve( 2 );
        int res[ 2 ] = { 0, 0 };
        MA_CUDA( cudaMemsetAsync( ctx->off_check.p, 0, 2 * sizeof( int ), ctx->stream ) );
        offsets_check_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>( ctx->read_off.p, n_reads, total, ctx->off_check.p );
        MA_CUDA( cudaGetLastError( ) );
        ctx->launches++;
        MA_CUDA( cudaMemcpyAsync( res, ctx->off_check.p, 2 * sizeof( int ), cudaMemcpyDeviceToHost, ctx->stream ) );
        MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
        maxL = res[ 0 ], bad = res[ 1 ] != 0;
    }
    else
        for( int64_t i = 0; i < n_reads && !bad; i++ )
        {
            const int64_t L = offsets[ i + 1 ] - offsets[ i ];
            if( L < 0 || L > 0x3fffffff || offsets[ i ] < 0 )
                bad = true;
            maxL = std::max<int>( maxL, (int)L );
        }
    if( bad )
    {
        if( ctx->upload_in_flight )
            cudaStreamSynchronize( ctx->copy_stream );
        ctx->upload_in_flight = false;
        ctx->n_reads = 0, ctx->stage_done = 0;
        throw std::runtime_error( "align_upload: bad offsets" );
    }
    ctx->n_reads = n_reads, ctx->max_read_len = maxL, ctx->stage_done = 0;
    ctx->reads_bytes = total;
    ctx->info.reserve( (size_t)n_reads + 1 );
    ctx->ctrl.reserve( 1 );
    if( ctx->upload_in_flight )
        return;
    ctx->reads.reserve( (size_t)ctx->reads_bytes + 16 );
    ctx->read_off.reserve( (size_t)n_reads + 1 );
    if( n_reads )
    {
        MA_CUDA( cudaMemcpyAsync( ctx->reads.p, reads, ctx->reads_bytes, cudaMemcpyHostToDevice, ctx->stream ) );
        MA_CUDA( cudaMemcpyAsync( ctx->read_off.p, offsets, ( n_reads + 1 ) * 8, cudaMemcpyHostToDevice,
                                  ctx->stream ) );
    }
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
}

extern "C" int ma_b200_align_upload( ma_b200_ctx* ctx, int64_t n_reads, const uint8_t* reads, const int64_t* offsets )
{
    MA_API_BEGIN
    align_upload_impl( ctx, n_reads, reads, offsets, false );
    MA_API_END
}

static void read_ctrl( ma_b200_ctx* ctx )
{
    MA_CUDA( cudaMemcpyAsync( &ctx->hctrl, ctx->ctrl.p, sizeof( PipeCtrl ), cudaMemcpyDeviceToHost, ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
}
template <typename T> static void zero_field( ma_b200_ctx* ctx, T PipeCtrl::*f )
{
    MA_CUDA( cudaMemsetAsync( (char*)ctx->ctrl.p + ( (size_t) & ( ( (PipeCtrl*)0 )->*f ) ), 0, sizeof( T ),
                              ctx->stream ) );
}
static float ev_ms( cudaEvent_t a, cudaEvent_t b )
{
    float ms = 0;
    cudaEventElapsedTime( &ms, a, b );
    return ms;
}

template <typename K> static int full_grid( ma_b200_ctx* ctx, K kernel, int threads, long long items, size_t smem = 0 )
{
    int perSm = 0;
    MA_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, kernel, threads, smem ) );
    if( perSm < 1 )
        perSm = 1;
    long long g = (long long)perSm * ctx->num_sms;
    g = std::min<long long>( g, ( items + threads - 1 ) / threads );
    return (int)std::max<long long>( g, 1 );
}

// DP over the planned tasks, bins decided on the device by nwbin_kernel. Two phases: the bins of ksw_qs_kernel, which
// may hand problems over into the bins of ksw_batch_kernel (PipeCtrl::bin_*), then those.
static void run_pipeline_dp( ma_b200_ctx* ctx, long long task_cap )
{
    NvtxRange nvtxDp( "banded DP (ksw_qs / ksw_tiny / ksw_batch kernels)" );
    static const int Ws[ 5 ] = { 128, 256, 512, 1024, 2048 };
    const KswScore score = make_score( ctx->params );
    ctx->task_out.reserve( (size_t)ctx->n_tasks + 1 );
    ctx->ksw_ctrl.reserve( 128 );
    // MA_B200_DP_BINS=1: time and cells of every DP launch (window class x kind) on stderr
    static const bool bBinStats = getenv( "MA_B200_DP_BINS" ) != nullptr;
    if( bBinStats && !ctx->binEv[ 0 ][ 0 ] )
        for( auto& e : ctx->binEv )
        {
            MA_CUDA( cudaEventCreate( &e[ 0 ] ) );
            MA_CUDA( cudaEventCreate( &e[ 1 ] ) );
        }
    int count0[ 16 ]; // bins of ksw_batch_kernel before ksw_qs_kernel adds to them
    for( int b = 0; b < 16; b++ )
        count0[ b ] = ctx->hctrl.bin_count[ b ];
    long long cigCap = std::max<long long>( ctx->task_cigar.cap, std::max<long long>( 12 * ctx->n_tasks, 1 << 16 ) );
    const size_t offCount = (size_t) & ( ( (PipeCtrl*)0 )->bin_count );
    for( int attempt = 0; attempt < 2; attempt++ )
    {
        ctx->task_cigar.reserve( (size_t)cigCap );
        cigCap = (long long)ctx->task_cigar.cap;
        MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p, 0, 128 * sizeof( unsigned long long ), ctx->stream ) );
        if( attempt > 0 )
            MA_CUDA( cudaMemcpyAsync( (char*)ctx->ctrl.p + offCount, count0, sizeof( count0 ), cudaMemcpyHostToDevice,
                                      ctx->stream ) );
        const long long budget = 48ll << 30; // traceback + cigar scratch of all resident warps (180 GB of HBM per GPU)
        KswBatchArgs A;
        A.tasks = ctx->tasks.p;
        A.seq = ctx->reads.p;
        A.pac = ctx->index.pac, A.fwd_len = ctx->index.fwd_len;
        A.out = ctx->task_out.p;
        A.cigar = ctx->task_cigar.p;
        A.cigar_cap = cigCap;
        A.cigar_cursor = ctx->ksw_ctrl.p;
        A.next = (int*)( ctx->ksw_ctrl.p + 16 );
        A.error = (int*)( ctx->ksw_ctrl.p + 32 );
        A.score = score;
        A.qsk[ 0 ] = ksw_qs_make_k( score, true ), A.qsk[ 1 ] = ksw_qs_make_k( score, false );
    A.bxk[ 0 ] = ksw_bx_make_k( score, true ), A.bxk[ 1 ] = ksw_bx_make_k( score, false );
    A.no_bx = no_bx_bits( );
        A.redo_count = ctx->ctrl.p->bin_count, A.redo_tb = ctx->ctrl.p->bin_tb, A.redo_cig = ctx->ctrl.p->bin_cig;
        A.redo_order = ctx->bin_order.p, A.redo_cap = task_cap, A.redo_n = nullptr;
        for( int phase = 0; phase < 2; phase++ )
        {
            const int b0 = phase == 0 ? MA_QS_BIN0 : 0, b1 = phase == 0 ? MA_TINY_BIN + 1 : 15;
            if( phase == 1 )
            {
                read_ctrl( ctx ); // synchronises: the bins now hold what ksw_qs_kernel handed over
                if( ctx->hctrl.bin_count[ 15 ] > 0 ) // (nwplan_kernel gives such sets up: MA_READ_EBAND)
                    throw std::runtime_error( "DP band wider than the largest supported window (internal error)" );
                for( int b = 0; b < 15; b++ )
                    if( ctx->hctrl.bin_count[ b ] > task_cap )
                        throw std::runtime_error( "pipeline DP: bin list overflow (internal error)" );
            }
            long long grids[ MA_NBINS ];
            size_t tbNeed = 0, csNeed = 0;
            for( int b = b0; b < b1; b++ )
            {
                grids[ b ] = 0;
                if( ctx->hctrl.bin_count[ b ] == 0 )
                    continue;
                if( b == MA_TINY_BIN )
                { // one thread per problem, no slabs
                    grids[ b ] = std::min<long long>( ( ctx->hctrl.bin_count[ b ] + 127 ) / 128, (long long)ctx->num_sms * 16 );
                    continue;
                }
                KswHostBin bin;
                bin.order.resize( ctx->hctrl.bin_count[ b ] ); // only its size is used
                bin.tb_stride = (long long)ctx->hctrl.bin_tb[ b ], bin.cig_stride = ctx->hctrl.bin_cig[ b ];
                if( phase == 0 )
                    bin.W = 0, bin.qs = b - MA_QS_BIN0 + 1;
                else
                    bin.W = Ws[ b / 3 ], bin.qs = 0;
                grids[ b ] = ksw_host_grid( ctx, bin, budget );
                const int wpc = bin.qs ? MA_QS_WARPS : ( bin.W == 1024 ? KswClass<1024>::WARPS : MA_KSW_WARPS );
                tbNeed = std::max<size_t>( tbNeed, (size_t)( grids[ b ] * wpc * bin.tb_stride ) );
                csNeed = std::max<size_t>( csNeed, (size_t)( grids[ b ] * wpc * bin.cig_stride ) );
            }
            ctx->ksw_tb.reserve( tbNeed + 256 );
            ctx->ksw_cigscratch.reserve( csNeed + 64 );
            A.tb = ctx->ksw_tb.p, A.cigscratch = ctx->ksw_cigscratch.p;
            for( int b = b0; b < b1; b++ )
            {
                if( ctx->hctrl.bin_count[ b ] == 0 )
                    continue;
                A.order = ctx->bin_order.p + (long long)b * task_cap;
                A.n = ctx->hctrl.bin_count[ b ];
                A.tb_stride = (long long)ctx->hctrl.bin_tb[ b ];
                A.cigscratch_stride = ctx->hctrl.bin_cig[ b ];
                A.cells_total = ctx->ksw_ctrl.p + 64 + b; // one counter per bin, summed below
                MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p + 16, 0, sizeof( unsigned long long ), ctx->stream ) );
                if( bBinStats )
                    MA_CUDA( cudaEventRecord( ctx->binEv[ b ][ 0 ], ctx->stream ) );
                if( b == MA_TINY_BIN )
                {
                    ksw_tiny_kernel<<<(unsigned)grids[ b ], 128, 0, ctx->stream>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches++;
                }
                else if( phase == 0 )
                    launch_ksw_qs( ctx, b - MA_QS_BIN0 + 1, A, grids[ b ] );
                else
                    switch( b / 3 )
                    {
                        case 0: launch_ksw_bin<128>( ctx, A, grids[ b ] ); break;
                        case 1: launch_ksw_bin<256>( ctx, A, grids[ b ] ); break;
                        case 2: launch_ksw_bin<512>( ctx, A, grids[ b ] ); break;
                        case 3: launch_ksw_bin<1024>( ctx, A, grids[ b ] ); break;
                        default: launch_ksw_bin<2048>( ctx, A, grids[ b ] ); break;
                    }
                if( bBinStats )
                    MA_CUDA( cudaEventRecord( ctx->binEv[ b ][ 1 ], ctx->stream ) );
            }
        }
        unsigned long long c[ 128 ];
        MA_CUDA( cudaMemcpyAsync( c, ctx->ksw_ctrl.p, sizeof( c ), cudaMemcpyDeviceToHost, ctx->stream ) );
        MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
        ctx->n_task_cigar = (int64_t)c[ 0 ];
        c[ 48 ] = 0;
        for( int b = 0; b < MA_NBINS; b++ )
            c[ 48 ] += c[ 64 + b ];
        if( bBinStats )
            for( int b = 0; b < MA_NBINS; b++ )
                if( b != 15 && ctx->hctrl.bin_count[ b ] > 0 && ctx->binEv[ b ][ 0 ] )
                {
                    const float ms = ev_ms( ctx->binEv[ b ][ 0 ], ctx->binEv[ b ][ 1 ] );
                    static const char* kKind[ 3 ] = { "all fields (exact)", "early-stop left", "early-stop right" };
                    if( b == MA_TINY_BIN )
                        fprintf( stderr, "ma_b200 dp bin tiny gap fills (one thread each): %d tasks, %llu cells, %.3f ms, %.1f GCUPS\n",
                                 ctx->hctrl.bin_count[ b ], c[ 64 + b ], ms, c[ 64 + b ] / ms / 1e6 );
                    else if( b >= MA_QS_BIN0 )
                        fprintf( stderr, "ma_b200 dp bin QS blocks=%d %s: %d tasks, %llu cells, %.3f ms, %.1f GCUPS\n",
                                 ( b - MA_QS_BIN0 ) / 2 + 1, ( b - MA_QS_BIN0 ) % 2 ? "right" : "left",
                                 ctx->hctrl.bin_count[ b ], c[ 64 + b ], ms, c[ 64 + b ] / ms / 1e6 );
                    else
                        fprintf( stderr, "ma_b200 dp bin W=%d %s: %d tasks, %llu cells, %.3f ms, %.1f GCUPS\n", Ws[ b / 3 ],
                                 kKind[ b % 3 ], ctx->hctrl.bin_count[ b ], c[ 64 + b ], ms, c[ 64 + b ] / ms / 1e6 );
                }
        ctx->ksw_cigar_used = c[ 48 ]; // reused as dp cell counter for the pipeline stats
        if( !(int)c[ 32 ] )
            return;
        cigCap = (long long)c[ 0 ] + 1024; // the cursor kept counting: exact size
    }
    throw std::runtime_error( "pipeline DP: cigar slab overflow after growing (internal error)" );
}

__global__ void stream_to_compute_kernel( )
{
}

extern "C" int ma_b200_align_run( ma_b200_ctx* ctx, int32_t upto_stage, int32_t keep_segments,
                                  ma_b200_align_stats* stats )
{
    MA_API_BEGIN
    if( !ctx->have_index )
    {
        ctx->err = "align_run: no index uploaded";
        return MA_B200_ESTATE;
    }
    if( upto_stage < 1 || upto_stage > 4 )
        throw std::runtime_error( "align_run: bad stage" );
    for( int i = 0; i < 8; i++ )
        if( !ctx->ev[ i ] )
            MA_CUDA( cudaEventCreate( &ctx->ev[ i ] ) );
    ma_b200_align_stats st;
    memset( &st, 0, sizeof( st ) );
    const int64_t launches0 = ctx->launches;
    const int n = (int)ctx->n_reads;
    st.n_reads = n;
    ctx->n_seeds = ctx->n_sets = ctx->n_set_seeds = ctx->n_tasks = ctx->n_runs = ctx->n_task_cigar = 0;
    ctx->stage_done = 0, ctx->compacted = false, ctx->n_reported = 0;
    cudaStream_t s = ctx->stream;
    NvtxRange nvtxRun( "ma_b200_align_run" );
    // The last operation of this stream was a device-to-host copy (the offsets check of the upload, the records of the
    // previous batch): an event recorded right behind it is queued on that copy engine, and with a second batch in
    // flight on the device it then waits for the OTHER batch's 200 MB record download (6 ms). An empty kernel moves the
    // stream to the compute engine first.
    stream_to_compute_kernel<<<1, 32, 0, s>>>( );
    ctx->launches++;
    MA_CUDA( cudaEventRecord( ctx->ev[ 0 ], s ) );
    static const bool bTraceRun = getenv( "MA_B200_E2E_TRACE" ) != nullptr;
    const auto tRun0 = std::chrono::steady_clock::now( );
    if( bTraceRun )
    { // how long the stream takes to reach the start of this run
        MA_CUDA( cudaEventSynchronize( ctx->ev[ 0 ] ) );
        fprintf( stderr, "ma_b200 e2e trace ctx %p: run start reached by the stream after %.2f ms\n", (void*)ctx,
                 std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now( ) - tRun0 ).count( ) );
    }
    if( n > 0 )
    {
        MA_CUDA( cudaMemsetAsync( ctx->ctrl.p, 0, sizeof( PipeCtrl ), s ) );
        // ---------------- stage 1: seeding
        nvtxRangePushA( "stage 1: BinarySeeding + ExtractSeeds" );
        const int maxL = ctx->max_read_len;
        const int list_cap = std::min( maxL + 2, 1024 ), fseg_cap = std::min( 2 * maxL + 8, 1 << 16 );
        const int SB = MA_SEED_BLOCK;
        int grid = full_grid( ctx, seed_kernel, SB, n );
        const size_t perThread = (size_t)( 2 * list_cap + 8 ) * sizeof( SegRec ) + (size_t)fseg_cap * sizeof( FSeg );
        grid = (int)std::max<size_t>( 1, std::min<size_t>( grid, ( (size_t)8 << 30 ) / ( perThread * SB ) ) );
        ctx->lists.reserve( (size_t)grid * SB * ( 2 * list_cap + 8 ) );
        ctx->fsegs.reserve( (size_t)grid * SB * fseg_cap );
        ctx->dbg_cap = keep_segments > 0 ? keep_segments : 0;
        if( ctx->dbg_cap )
        {
            ctx->dbg_segs.reserve( (size_t)n * ctx->dbg_cap );
            ctx->dbg_nsegs.reserve( (size_t)n );
        }
        long long seedCap = std::max<long long>( ctx->seeds.cap, (long long)n * std::max( 8, maxL / 16 ) + 1024 );
        for( int attempt = 0;; attempt++ )
        {
            ctx->seeds.reserve( (size_t)seedCap );
            seedCap = (long long)ctx->seeds.cap;
            SeedKernelArgs A;
            A.I = ctx->index, A.P = make_seed_params( ctx->params );
            A.reads = ctx->reads.p, A.read_off = ctx->read_off.p, A.n_reads = n, A.info = ctx->info.p;
            A.seeds = ctx->seeds.p, A.seed_cap = seedCap;
            A.lists = ctx->lists.p, A.list_cap = list_cap, A.fsegs = ctx->fsegs.p, A.fseg_cap = fseg_cap;
            A.dbg_segs = ctx->dbg_cap ? ctx->dbg_segs.p : nullptr;
            A.dbg_nsegs = ctx->dbg_cap ? ctx->dbg_nsegs.p : nullptr, A.dbg_cap = ctx->dbg_cap;
            A.ctrl = ctx->ctrl.p;
            A.reads_ready = ctx->upload_in_flight ? ctx->reads_ready.p : nullptr;
            seed_kernel<<<grid, SB, 0, s>>>( A );
            MA_CUDA( cudaGetLastError( ) );
            ctx->launches++;
            read_ctrl( ctx );
            if( (long long)ctx->hctrl.seed_cursor <= seedCap )
                break;
            if( attempt > 0 )
                throw std::runtime_error( "seeding: seed slab overflow after growing (internal error)" );
            seedCap = (long long)ctx->hctrl.seed_cursor + 1024;
            MA_CUDA( cudaMemsetAsync( ctx->ctrl.p, 0, sizeof( PipeCtrl ), s ) );
        }
        ctx->n_seeds = (int64_t)ctx->hctrl.seed_cursor;
        st.n_ext = (int64_t)ctx->hctrl.n_ext, st.n_lookup = (int64_t)ctx->hctrl.n_lookup, st.n_dropped = (int64_t)ctx->hctrl.n_dropped;
        MA_CUDA( cudaEventRecord( ctx->ev[ 1 ], s ) );
        if( ctx->upload_in_flight )
            MA_CUDA( cudaStreamWaitEvent( s, ctx->ev_copy[ 1 ], 0 ) );
        if( ctx->n_seeds > 0 )
        {
            LocateArgs A{ ctx->index, ctx->seeds.p, ctx->n_seeds, ctx->read_off.p, ctx->ctrl.p };
            locate_kernel<<<full_grid( ctx, locate_kernel, 256, ctx->n_seeds ), 256, 0, s>>>( A );
            MA_CUDA( cudaGetLastError( ) );
            ctx->launches++;
        }
        MA_CUDA( cudaEventRecord( ctx->ev[ 2 ], s ) );
        nvtxRangePop( );
        ctx->stage_done = 1;
        // ---------------- stage 2: SoC + harmonization
        if( upto_stage >= 2 )
        {
            NvtxRange nvtxStage( "stage 2: StripOfConsideration + Harmonization" );
            const size_t scratchCap = ( ( (size_t)ctx->n_seeds * ( 340 + sizeof( DSeed ) ) + (size_t)n * 400 + 4096 ) + 255 ) & ~(size_t)255;
            ctx->harm_scratch.reserve( scratchCap );
            long long setSeedCap = std::max<long long>( ctx->set_seeds.cap, 2 * ctx->n_seeds + 1024 );
            long long setCap = std::max<long long>( ctx->sets.cap, 2ll * n + 1024 );
            for( int attempt = 0;; attempt++ )
            {
                ctx->set_seeds.reserve( (size_t)setSeedCap );
                ctx->sets.reserve( (size_t)setCap );
                setSeedCap = (long long)ctx->set_seeds.cap, setCap = (long long)ctx->sets.cap;
                SocHarmArgs A;
                A.I = ctx->index, A.P = make_harm_params( ctx->params ), A.read_off = ctx->read_off.p, A.n_reads = n;
                A.info = ctx->info.p, A.seeds = ctx->seeds.p;
                A.scratch = ctx->harm_scratch.p, A.scratch_cap = ctx->harm_scratch.cap;
                A.set_seeds = ctx->set_seeds.p, A.set_seed_cap = setSeedCap, A.sets = ctx->sets.p, A.set_cap = setCap;
                A.srand_base = ctx->params.srand_base, A.ctrl = ctx->ctrl.p;
                static const bool bOneKernel = getenv( "MA_B200_SOC_ONE_KERNEL" ) && atoi( getenv( "MA_B200_SOC_ONE_KERNEL" ) );
                if( bOneKernel )
                {
                    A.soc_scratch = nullptr, A.soc_nmax = nullptr;
                    socharm_kernel<<<full_grid( ctx, socharm_kernel, MA_SOC_BLOCK, n ), MA_SOC_BLOCK, 0, s>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches++;
                }
                else
                {
                    ctx->soc_scratch.reserve( (size_t)n + 1 );
                    ctx->soc_nmax.reserve( (size_t)n + 1 );
                    A.soc_scratch = ctx->soc_scratch.p, A.soc_nmax = ctx->soc_nmax.p;
                    socbuild_kernel<<<full_grid( ctx, socbuild_kernel, MA_SOC_BLOCK, n ), MA_SOC_BLOCK, 0, s>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    harmonize_kernel<<<full_grid( ctx, harmonize_kernel, MA_SOC_BLOCK, n ), MA_SOC_BLOCK, 0, s>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches += 2;
                }
                read_ctrl( ctx );
                if( ctx->hctrl.scratch_cursor > ctx->harm_scratch.cap )
                    throw std::runtime_error( "harmonization: scratch arena too small (internal error)" );
                if( (long long)ctx->hctrl.set_seed_cursor <= setSeedCap && (long long)ctx->hctrl.set_cursor <= setCap )
                    break;
                if( attempt > 0 )
                    throw std::runtime_error( "harmonization: slab overflow after growing (internal error)" );
                setSeedCap = (long long)ctx->hctrl.set_seed_cursor + 1024;
                setCap = (long long)ctx->hctrl.set_cursor + 1024;
                zero_field( ctx, &PipeCtrl::set_seed_cursor );
                zero_field( ctx, &PipeCtrl::set_cursor );
                zero_field( ctx, &PipeCtrl::scratch_cursor );
                zero_field( ctx, &PipeCtrl::next_read2 );
                zero_field( ctx, &PipeCtrl::next_read3 );
            }
            ctx->n_sets = (int64_t)ctx->hctrl.set_cursor, ctx->n_set_seeds = (int64_t)ctx->hctrl.set_seed_cursor;
            ctx->stage_done = 2;
        }
        MA_CUDA( cudaEventRecord( ctx->ev[ 3 ], s ) );
        // ---------------- stage 3: NW
        if( upto_stage >= 3 )
        {
            NvtxRange nvtxStage( "stage 3 + 4: NeedlemanWunsch, MappingQuality, PairedReads" );
            const int nSets = (int)ctx->n_sets;
            long long taskCap = std::max<long long>( ( ctx->bin_order.cap / MA_NBINS ), 3ll * nSets + 1024 );
            if( nSets > 0 )
                for( int attempt = 0;; attempt++ )
                {
                    ctx->tasks.reserve( (size_t)taskCap );
                    ctx->bin_order.reserve( (size_t)taskCap * MA_NBINS );
                    NwPlanArgs A;
                    A.I = ctx->index, A.P = make_nw_params( ctx->params ), A.read_off = ctx->read_off.p;
                    A.sets = ctx->sets.p, A.n_sets = nSets, A.set_seeds = ctx->set_seeds.p;
                    A.tasks = ctx->tasks.p, A.task_cap = taskCap, A.bin_order = ctx->bin_order.p, A.ctrl = ctx->ctrl.p;
                    A.info = ctx->info.p;
                    nwplan_kernel<<<full_grid( ctx, nwplan_kernel, 128, nSets ), 128, 0, s>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches++;
                    read_ctrl( ctx );
                    if( (long long)ctx->hctrl.task_cursor <= taskCap )
                        break;
                    if( attempt > 0 )
                        throw std::runtime_error( "NW planning: task slab overflow after growing (internal error)" );
                    taskCap = (long long)ctx->hctrl.task_cursor + 1024;
                    zero_field( ctx, &PipeCtrl::task_cursor );
                }
            ctx->n_tasks = nSets > 0 ? (int64_t)ctx->hctrl.task_cursor : 0;
            if( ctx->n_tasks > 0 )
            { // window bins of the tasks
                NwBinArgs B{ ctx->tasks.p, (int)ctx->n_tasks, taskCap, ctx->bin_order.p, ctx->ctrl.p, make_score( ctx->params ),
                             use_qs( ) ? 1 : 0, use_tiny( ) ? 1 : 0 };
                nwbin_kernel<<<full_grid( ctx, nwbin_kernel, 256, ctx->n_tasks ), 256, 0, s>>>( B );
                MA_CUDA( cudaGetLastError( ) );
                ctx->launches++;
                read_ctrl( ctx );
            }
            MA_CUDA( cudaEventRecord( ctx->ev[ 4 ], s ) );
            if( ctx->n_tasks > 0 )
                run_pipeline_dp( ctx, taskCap );
            st.dp_cells = ctx->n_tasks > 0 ? (int64_t)ctx->ksw_cigar_used : 0;
            MA_CUDA( cudaEventRecord( ctx->ev[ 5 ], s ) );
            ctx->alns.reserve( (size_t)nSets + 1 );
            if( nSets > 0 )
            {
                const int runScratchCap = 2 * ctx->max_read_len + 4096;
                int grid = full_grid( ctx, nwasm_kernel, 128, nSets );
                // per-thread run scratch of 2 L + 4096 words: one 100 kbp read in the batch must not turn the full
                // occupancy grid into a 100 GB allocation, so the grid is capped against a byte budget like stage 1's
                grid = (int)std::max<size_t>( 1, std::min<size_t>( (size_t)grid, ( (size_t)8 << 30 ) / ( (size_t)128 * runScratchCap * sizeof( unsigned int ) ) ) );
                ctx->run_scratch.reserve( (size_t)grid * 128 * runScratchCap );
                long long runCap = std::max<long long>( ctx->runs.cap, 16ll * nSets + 4096 );
                for( int attempt = 0;; attempt++ )
                {
                    ctx->runs.reserve( (size_t)runCap );
                    runCap = (long long)ctx->runs.cap;
                    NwAsmArgs A;
                    A.I = ctx->index, A.P = make_nw_params( ctx->params ), A.reads = ctx->reads.p;
                    A.read_off = ctx->read_off.p, A.sets = ctx->sets.p, A.n_sets = nSets;
                    A.set_seeds = ctx->set_seeds.p, A.res = ctx->task_out.p, A.cigar = ctx->task_cigar.p;
                    A.alns = ctx->alns.p, A.runs = ctx->runs.p, A.run_cap = runCap;
                    A.run_scratch = ctx->run_scratch.p, A.run_scratch_cap = runScratchCap, A.ctrl = ctx->ctrl.p;
                    nwasm_kernel<<<grid, 128, 0, s>>>( A );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches++;
                    read_ctrl( ctx );
                    if( ctx->hctrl.overflow_runs )
                        throw std::runtime_error( "NW assembly: run list scratch too small (internal error)" );
                    if( (long long)ctx->hctrl.run_cursor <= runCap )
                        break;
                    if( attempt > 0 )
                        throw std::runtime_error( "NW assembly: run slab overflow after growing (internal error)" );
                    runCap = (long long)ctx->hctrl.run_cursor + 1024;
                    zero_field( ctx, &PipeCtrl::run_cursor );
                    zero_field( ctx, &PipeCtrl::next_set );
                }
                ctx->n_runs = (int64_t)ctx->hctrl.run_cursor;
                if( ctx->early_runs && ctx->n_runs > 0 && ctx->n_runs <= ctx->early_runs_cap )
                { // the run words are final: their download runs under the kernels that follow
                    MA_CUDA( cudaEventRecord( ctx->ev_copy[ 2 ], s ) );
                    MA_CUDA( cudaStreamWaitEvent( ctx->copy_stream, ctx->ev_copy[ 2 ], 0 ) );
                    // (in pieces: the control read-backs of the kernels that follow share the copy engine)
                    const size_t bytes = (size_t)ctx->n_runs * sizeof( unsigned int ), piece = 4u << 20;
                    for( size_t o = 0; o < bytes; o += piece )
                        MA_CUDA( cudaMemcpyAsync( (char*)ctx->early_runs + o, (const char*)ctx->runs.p + o,
                                                  std::min( piece, bytes - o ), cudaMemcpyDeviceToHost, ctx->copy_stream ) );
                    ctx->early_runs_done = true;
                }
                AlnSortArgs B{ ctx->info.p, n, ctx->alns.p, ctx->ctrl.p };
                alnsort_kernel<<<full_grid( ctx, alnsort_kernel, 128, n ), 128, 0, s>>>( B );
                MA_CUDA( cudaGetLastError( ) );
                ctx->launches++;
                if( upto_stage >= 4 )
                { // ---------------- stage 4: MappingQuality (+ PairedReads)
                    const ma_b200_params& p = ctx->params;
                    MapqArgs M;
                    M.P = MapqParams{ p.match, p.report_n, p.min_alignment_score, p.max_supplementary_per_prim,
                                      p.max_overlap_supplementary, p.paired_mean, p.paired_std, p.paired_bonus };
                    M.info = ctx->info.p, M.read_off = ctx->read_off.p, M.n_reads = n, M.alns = ctx->alns.p;
                    M.runs = ctx->runs.p, M.ref_len = ctx->index.ref_len, M.ctrl = ctx->ctrl.p;
                    M.pair_sc = nullptr, M.pair_meta = nullptr, M.pair_cap = 0;
                    mapq_kernel<<<full_grid( ctx, mapq_kernel, 128, n ), 128, 0, s>>>( M );
                    MA_CUDA( cudaGetLastError( ) );
                    ctx->launches++;
                    if( p.use_paired_reads )
                    {
                        read_ctrl( ctx );
                        const int most = std::max( 1, ctx->hctrl.max_reported );
                        const int grid = full_grid( ctx, pair_kernel, 128, ( n + 1 ) / 2 );
                        M.pair_cap = most * most;
                        ctx->pair_sc.reserve( (size_t)grid * 128 * M.pair_cap );
                        ctx->pair_meta.reserve( (size_t)grid * 128 * 2 * M.pair_cap );
                        M.pair_sc = ctx->pair_sc.p, M.pair_meta = ctx->pair_meta.p;
                        pair_kernel<<<grid, 128, 0, s>>>( M );
                        MA_CUDA( cudaGetLastError( ) );
                        ctx->launches++;
                        read_ctrl( ctx );
                        if( ctx->hctrl.overflow_pair )
                            throw std::runtime_error( "PairedReads: no candidate pair for two aligned mates" );
                    }
                    if( ctx->reported_only )
                    { // the writer's records of every read next to each other (ma_b200_set_reported_only)
                        ctx->info_out.reserve( (size_t)n + 1 );
                        ctx->alns_out.reserve( (size_t)nSets + 1 );
                        zero_field( ctx, &PipeCtrl::reported_cursor );
                        CompactArgs C{ ctx->info.p, ctx->info_out.p, n, ctx->alns.p, ctx->alns_out.p, ctx->ctrl.p };
                        compact_reported_kernel<<<full_grid( ctx, compact_reported_kernel, 128, n ), 128, 0, s>>>( C );
                        MA_CUDA( cudaGetLastError( ) );
                        ctx->launches++;
                        ctx->compacted = true;
                    }
                }
            }
            ctx->stage_done = upto_stage;
        }
        else
        {
            MA_CUDA( cudaEventRecord( ctx->ev[ 4 ], s ) );
            MA_CUDA( cudaEventRecord( ctx->ev[ 5 ], s ) );
        }
        read_ctrl( ctx );
        st.n_invpsi = (int64_t)ctx->hctrl.n_invpsi;
    }
    else
    {
        for( int i = 1; i < 6; i++ )
            MA_CUDA( cudaEventRecord( ctx->ev[ i ], s ) );
        ctx->stage_done = upto_stage; // an empty batch has (empty) results for every stage
    }
    MA_CUDA( cudaEventRecord( ctx->ev[ 6 ], s ) );
    MA_CUDA( cudaEventSynchronize( ctx->ev[ 6 ] ) );
    if( bTraceRun )
        fprintf( stderr, "ma_b200 e2e trace ctx %p: run wall %.2f ms, event span %.2f ms\n", (void*)ctx,
                 std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now( ) - tRun0 ).count( ),
                 ev_ms( ctx->ev[ 0 ], ctx->ev[ 6 ] ) );
    st.n_seeds = ctx->n_seeds, st.n_sets = ctx->n_sets, st.n_set_seeds = ctx->n_set_seeds, st.n_tasks = ctx->n_tasks;
    st.n_runs = ctx->n_runs, st.n_cigar_words = ctx->n_task_cigar;
    st.ms_seed = ev_ms( ctx->ev[ 0 ], ctx->ev[ 1 ] ), st.ms_locate = ev_ms( ctx->ev[ 1 ], ctx->ev[ 2 ] );
    st.ms_socharm = ev_ms( ctx->ev[ 2 ], ctx->ev[ 3 ] ), st.ms_plan = ev_ms( ctx->ev[ 3 ], ctx->ev[ 4 ] );
    st.ms_dp = ev_ms( ctx->ev[ 4 ], ctx->ev[ 5 ] ), st.ms_assemble = ev_ms( ctx->ev[ 5 ], ctx->ev[ 6 ] );
    st.ms_total = ev_ms( ctx->ev[ 0 ], ctx->ev[ 6 ] );
    st.launches = (int)( ctx->launches - launches0 );
    st.n_failed = n > 0 ? ctx->hctrl.n_failed : 0;
    ctx->n_reported = ctx->compacted ? (int64_t)ctx->hctrl.reported_cursor : 0;
    st.n_reported = ctx->n_reported;
    if( stats )
        *stats = st;
    MA_API_END
}

extern "C" int ma_b200_align_download_info( ma_b200_ctx* ctx, ma_b200_read_info* info )
{
    MA_API_BEGIN
    static_assert( sizeof( ma_b200_read_info ) == sizeof( ReadInfo ), "read info layout" );
    if( ctx->stage_done < 1 )
    {
        ctx->err = "nothing to download";
        return MA_B200_ESTATE;
    }
    if( ctx->n_reads )
        MA_CUDA( cudaMemcpyAsync( info, ctx->info.p, ctx->n_reads * sizeof( ReadInfo ), cudaMemcpyDeviceToHost,
                                  ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

extern "C" int ma_b200_align_download_segments( ma_b200_ctx* ctx, ma_b200_segment* segs, int32_t* n_segs )
{
    MA_API_BEGIN
    static_assert( sizeof( ma_b200_segment ) == sizeof( SegRec ), "segment layout" );
    if( ctx->stage_done < 1 || ctx->dbg_cap == 0 )
    {
        ctx->err = "segments were not kept (keep_segments == 0)";
        return MA_B200_ESTATE;
    }
    if( ctx->n_reads )
    {
        MA_CUDA( cudaMemcpyAsync( segs, ctx->dbg_segs.p, ctx->n_reads * ctx->dbg_cap * sizeof( SegRec ),
                                  cudaMemcpyDeviceToHost, ctx->stream ) );
        MA_CUDA( cudaMemcpyAsync( n_segs, ctx->dbg_nsegs.p, ctx->n_reads * sizeof( int ), cudaMemcpyDeviceToHost,
                                  ctx->stream ) );
    }
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

extern "C" int ma_b200_align_download_seeds( ma_b200_ctx* ctx, ma_b200_seed* seeds, int64_t cap )
{
    MA_API_BEGIN
    static_assert( sizeof( ma_b200_seed ) == sizeof( DSeed ), "seed layout" );
    if( ctx->stage_done != 1 )
    {
        ctx->err = "seeds are only available right after a run with upto_stage == MA_B200_STAGE_SEEDS";
        return MA_B200_ESTATE;
    }
    if( cap < ctx->n_seeds )
    {
        ctx->err = "seed buffer too small";
        return MA_B200_ENOMEM;
    }
    if( ctx->n_seeds )
        MA_CUDA( cudaMemcpyAsync( seeds, ctx->seeds.p, ctx->n_seeds * sizeof( DSeed ), cudaMemcpyDeviceToHost,
                                  ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

extern "C" int ma_b200_align_download_sets( ma_b200_ctx* ctx, ma_b200_seed_set* sets, int64_t cap_sets,
                                            ma_b200_seed* seeds, int64_t cap_seeds )
{
    MA_API_BEGIN
    static_assert( sizeof( ma_b200_seed_set ) == sizeof( SetHeader ), "set layout" );
    if( ctx->stage_done < 2 )
    {
        ctx->err = "no seed sets computed";
        return MA_B200_ESTATE;
    }
    if( cap_sets < ctx->n_sets || cap_seeds < ctx->n_set_seeds )
    {
        ctx->err = "set buffers too small";
        return MA_B200_ENOMEM;
    }
    if( ctx->n_sets )
        MA_CUDA( cudaMemcpyAsync( sets, ctx->sets.p, ctx->n_sets * sizeof( SetHeader ), cudaMemcpyDeviceToHost,
                                  ctx->stream ) );
    if( ctx->n_set_seeds )
        MA_CUDA( cudaMemcpyAsync( seeds, ctx->set_seeds.p, ctx->n_set_seeds * sizeof( DSeed ), cudaMemcpyDeviceToHost,
                                  ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

extern "C" int ma_b200_align_download( ma_b200_ctx* ctx, ma_b200_read_info* info, ma_b200_alignment* alns,
                                       int64_t cap_alns, uint32_t* runs, int64_t cap_runs )
{
    NvtxRange nvtxCall( "ma_b200_align_download" );
    MA_API_BEGIN
    static_assert( sizeof( ma_b200_alignment ) == sizeof( DAln ), "alignment layout" );
    if( ctx->stage_done < 3 )
    {
        ctx->err = "no alignments computed";
        return MA_B200_ESTATE;
    }
    const bool compact = ctx->compacted && ctx->stage_done == 4;
    const int64_t nAlns = compact ? ctx->n_reported : ctx->n_sets;
    if( cap_alns < nAlns || cap_runs < ctx->n_runs )
    {
        ctx->err = "alignment buffers too small";
        return MA_B200_ENOMEM;
    }
    // in pieces: with two batches in flight on a device (sibling contexts) the small control read-backs of the OTHER
    // batch's stages share the copy engine and would otherwise wait for a whole 200 MB record download
    auto down = [ & ]( void* dst, const void* src, size_t bytes ) {
        const size_t piece = 4u << 20;
        for( size_t o = 0; o < bytes; o += piece )
            MA_CUDA( cudaMemcpyAsync( (char*)dst + o, (const char*)src + o, std::min( piece, bytes - o ), cudaMemcpyDeviceToHost,
                                      ctx->stream ) );
    };
    if( ctx->n_reads && info )
        down( info, compact ? ctx->info_out.p : ctx->info.p, ctx->n_reads * sizeof( ReadInfo ) );
    if( nAlns )
        down( alns, compact ? ctx->alns_out.p : ctx->alns.p, nAlns * sizeof( DAln ) );
    if( ctx->n_runs && !( ctx->early_runs_done && runs == ctx->early_runs ) )
        down( runs, ctx->runs.p, ctx->n_runs * sizeof( unsigned int ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    if( ctx->early_runs_done )
        MA_CUDA( cudaStreamSynchronize( ctx->copy_stream ) );
    MA_API_END
}

extern "C" int ma_b200_set_reported_only( ma_b200_ctx* ctx, int32_t on )
{
    if( !ctx )
        return MA_B200_EINVAL;
    ctx->reported_only = on != 0;
    if( ctx->shadow )
        ctx->shadow->reported_only = ctx->reported_only;
    return MA_B200_OK;
}

extern "C" int ma_b200_set_batch_split( ma_b200_ctx* ctx, int64_t reads_per_subbatch )
{
    if( !ctx || reads_per_subbatch < 0 )
        return MA_B200_EINVAL;
    ctx->batch_split = reads_per_subbatch;
    return MA_B200_OK;
}

namespace
{
// Pipelined form of ma_b200_align_batch: the batch is cut into sub-batches that alternate between two sets of device
// slabs (the context and its shadow, one host thread and one stream each), so that the host<->device copies of one
// sub-batch run under the kernels of the other. Results are identical to the one-shot form: RANSAC streams are
// seeded by the global read index, outputs are written at their final offsets.
struct BatchPipe
{
    int64_t n_reads, n_sub;
    const uint8_t* reads;
    const int64_t* offsets;
    ma_b200_read_info* info;
    ma_b200_alignment* alns;
    int64_t cap_alns;
    uint32_t* runs;
    int64_t cap_runs;
    int64_t split;
    uint32_t srand_base;
    std::mutex mtx;
    std::condition_variable cv;
    std::vector<int64_t> nAlns, nRuns; // per sub-batch, -1 until its kernels are done
    std::vector<ma_b200_align_stats> st;
    int rc = 0;
    std::string err;

    void fail( int code, const std::string& msg )
    {
        std::lock_guard<std::mutex> g( mtx );
        if( !rc )
            rc = code, err = msg;
        cv.notify_all( );
    }
    void work( ma_b200_ctx* c, int64_t first )
    {
        cudaSetDevice( c->device );
        std::vector<int64_t> off;
        for( int64_t k = first; k < n_sub; k += 2 )
        {
            {
                std::lock_guard<std::mutex> g( mtx );
                if( rc )
                    return;
            }
            const int64_t r0 = k * split, r1 = std::min( n_reads, r0 + split ), n = r1 - r0;
            off.resize( (size_t)n + 1 );
            for( int64_t i = 0; i <= n; i++ )
                off[ i ] = offsets[ r0 + i ] - offsets[ r0 ];
            c->params.srand_base = srand_base + (uint32_t)r0;
            int e = ma_b200_align_upload( c, n, reads + offsets[ r0 ], off.data( ) );
            if( !e )
                e = ma_b200_align_run( c, MA_B200_STAGE_MAPQ, 0, &st[ k ] );
            if( e )
                return fail( e, c->err );
            int64_t a0 = 0, u0 = 0;
            {
                std::unique_lock<std::mutex> g( mtx );
                nAlns[ k ] = c->compacted ? c->n_reported : c->n_sets, nRuns[ k ] = c->n_runs;
                cv.notify_all( );
                cv.wait( g, [ & ] {
                    if( rc )
                        return true;
                    for( int64_t j = 0; j < k; j++ )
                        if( nAlns[ j ] < 0 )
                            return false;
                    return true;
                } );
                if( rc )
                    return;
                for( int64_t j = 0; j < k; j++ )
                    a0 += nAlns[ j ], u0 += nRuns[ j ];
            }
            const int64_t nOut = c->compacted ? c->n_reported : c->n_sets;
            if( a0 + nOut > cap_alns || u0 + c->n_runs > cap_runs )
                return fail( MA_B200_ENOMEM, "alignment buffers too small" );
            e = ma_b200_align_download( c, info + r0, alns + a0, cap_alns - a0, runs + u0, cap_runs - u0 );
            if( e )
                return fail( e, c->err );
            // sub-batch-relative indices -> batch-relative (seed_off keeps pointing into the device slab)
            for( int64_t i = 0; i < n; i++ )
                info[ r0 + i ].set_off += (int32_t)a0;
            for( int64_t j = 0; j < nOut; j++ )
                alns[ a0 + j ].read += (int32_t)r0, alns[ a0 + j ].run_off += u0;
        }
    }
};
} // namespace

extern "C" int ma_b200_align_batch( ma_b200_ctx* ctx, int64_t n_reads, const uint8_t* reads, const int64_t* offsets,
                                    ma_b200_read_info* info, ma_b200_alignment* alns, int64_t cap_alns,
                                    uint32_t* runs, int64_t cap_runs, ma_b200_align_stats* stats )
{
    NvtxRange nvtxCall( "ma_b200_align_batch" );
    if( !ctx )
        return MA_B200_EINVAL;
    if( ctx->batch_split <= 0 || n_reads < 2 * ( ctx->batch_split + ( ctx->batch_split & 1 ) ) || !ctx->have_index || !reads || !offsets || !info || !alns || !runs )
    { // one shot (also the path that reports argument errors): the upload runs under the seeding kernel, the download
      // of the run words under the kernels of stage 4
        int rc = MA_B200_OK;
        // MA_B200_E2E_TRACE=1: host time stamps of the phases of every call (ms since the first call) on stderr
        static const bool bTrace = getenv( "MA_B200_E2E_TRACE" ) != nullptr;
        static const auto t00 = std::chrono::steady_clock::now( );
        auto now = [ & ]( ) { return std::chrono::duration<double, std::milli>( std::chrono::steady_clock::now( ) - t00 ).count( ); };
        const double tEnter = now( );
        try
        {
            MA_CUDA( cudaSetDevice( ctx->device ) );
            ensure_copy_stream( ctx );
            align_upload_impl( ctx, n_reads, reads, offsets, true );
        }
        catch( const ma::CudaError& e )
        {
            ctx->err = e.msg, rc = MA_B200_ECUDA;
        }
        catch( const std::exception& e )
        {
            ctx->err = e.what( ), rc = MA_B200_EINVAL;
        }
        ctx->early_runs = runs, ctx->early_runs_cap = runs ? cap_runs : 0, ctx->early_runs_done = false;
        double tUp = 0, tLock = 0, tRun = 0;
        if( !rc )
        { // One batch at a time computes on a device: when two host threads keep two batches in flight on sibling
          // contexts, the kernels of the two would otherwise interleave, both batches would finish together and their
          // downloads would find the GPU idle. With the turnstile the upload of a batch (enqueued above) and the
          // download of its records (below) run under the OTHER batch's kernels.
            tUp = now( );
            std::lock_guard<std::mutex> turn( device_turnstile( ctx->device ) );
            tLock = now( );
            rc = ma_b200_align_run( ctx, MA_B200_STAGE_MAPQ, 0, stats );
            tRun = now( );
        }
        if( ctx->upload_in_flight )
        { // whatever happened above, nothing of the caller's buffers may still be in flight when this call returns
            cudaStreamSynchronize( ctx->copy_stream );
            ctx->upload_in_flight = false;
        }
        ctx->early_runs = nullptr;
        if( !rc )
            rc = ma_b200_align_download( ctx, info, alns, cap_alns, runs, cap_runs );
        else if( ctx->early_runs_done )
            cudaStreamSynchronize( ctx->copy_stream );
        ctx->early_runs_done = false;
        if( bTrace )
            fprintf( stderr, "ma_b200 e2e trace ctx %p: enter %.2f upload-call %.2f lock-wait %.2f run %.2f download %.2f (end %.2f)\n",
                     (void*)ctx, tEnter, tUp - tEnter, tLock - tUp, tRun - tLock, now( ) - tRun, now( ) );
        return rc;
    }
    if( !ctx->shadow )
    {
        const int rc = ma_b200_create( ctx->device, &ctx->shadow );
        if( rc )
        {
            ctx->err = "align_batch: cannot create the second pipeline context";
            return rc;
        }
    }
    ma_b200_ctx* sh = ctx->shadow;
    const ma_b200_params saved = ctx->params;
    sh->params = ctx->params, sh->reported_only = ctx->reported_only;
    sh->index = ctx->index, sh->have_index = true; // a view: the index slabs stay owned by ctx
    const int64_t l0 = ctx->launches, l1 = sh->launches;
    BatchPipe P;
    P.n_reads = n_reads, P.split = ctx->batch_split + ( ctx->batch_split & 1 ), P.n_sub = ( n_reads + P.split - 1 ) / P.split;
    P.reads = reads, P.offsets = offsets, P.info = info, P.alns = alns, P.cap_alns = cap_alns, P.runs = runs;
    P.cap_runs = cap_runs, P.srand_base = saved.srand_base;
    P.nAlns.assign( (size_t)P.n_sub, -1 ), P.nRuns.assign( (size_t)P.n_sub, -1 );
    P.st.assign( (size_t)P.n_sub, ma_b200_align_stats{ } );
    std::thread other( [ & ] { P.work( sh, 1 ); } );
    P.work( ctx, 0 );
    other.join( );
    cudaSetDevice( ctx->device );
    ctx->params = saved;
    ctx->stage_done = 0; // the device slabs hold the last sub-batches only: staged downloads need a staged run
    sh->stage_done = 0;
    if( P.rc )
    {
        ctx->err = P.err;
        return P.rc;
    }
    if( stats )
    {
        ma_b200_align_stats t{ };
        for( const auto& s : P.st )
        {
            t.n_reads += s.n_reads, t.n_seeds += s.n_seeds, t.n_sets += s.n_sets, t.n_set_seeds += s.n_set_seeds;
            t.n_tasks += s.n_tasks, t.n_runs += s.n_runs, t.n_cigar_words += s.n_cigar_words, t.n_ext += s.n_ext;
            t.n_invpsi += s.n_invpsi, t.n_dropped += s.n_dropped, t.dp_cells += s.dp_cells, t.n_lookup += s.n_lookup;
            t.n_failed += s.n_failed, t.n_reported += s.n_reported;
            t.ms_seed += s.ms_seed, t.ms_locate += s.ms_locate, t.ms_socharm += s.ms_socharm, t.ms_plan += s.ms_plan;
            t.ms_dp += s.ms_dp, t.ms_assemble += s.ms_assemble, t.ms_total += s.ms_total;
        }
        t.launches = (int32_t)( ( ctx->launches - l0 ) + ( sh->launches - l1 ) );
        *stats = t;
    }
    ctx->launches += sh->launches - l1; // ma_b200_launch_count counts both pipelines
    return MA_B200_OK;
}

extern "C" int ma_b200_paired_reads_host( const ma_b200_params* params, int64_t ref_len, ma_b200_alignment* mate1,
                                          int32_t n1, int64_t qlen1, ma_b200_alignment* mate2, int32_t n2, int64_t qlen2,
                                          const uint32_t* runs )
{
    if( !params || n1 < 0 || n2 < 0 || ( n1 && !mate1 ) || ( n2 && !mate2 ) || n1 > 0x7fff || n2 > 0x7fff )
        return MA_B200_EINVAL;
    const ma_b200_params& p = *params;
    const MapqParams M{ p.match, p.report_n, p.min_alignment_score, p.max_supplementary_per_prim,
                        p.max_overlap_supplementary, p.paired_mean, p.paired_std, p.paired_bonus };
    const int cap = std::max( 1, n1 * n2 );
    std::vector<int> ord1( (size_t)n1 + 1 ), ord2( (size_t)n2 + 1 ), meta( 2 * (size_t)cap );
    std::vector<long long> sc( (size_t)cap );
    const int n = paired_reads_pair( M, ref_len, reinterpret_cast<DAln*>( mate1 ), n1, qlen1, reinterpret_cast<DAln*>( mate2 ),
                                     n2, qlen2, runs, ord1.data( ), ord2.data( ), sc.data( ), meta.data( ), cap );
    return n < 0 ? MA_B200_EINVAL : n;
}

// ------------------------------------------------------------------------------------------------ roofline probe
extern "C" int ma_b200_gather_probe( ma_b200_ctx* ctx, int64_t buffer_bytes, double* gbs )
{
    MA_API_BEGIN
    if( buffer_bytes < 4096 || !gbs )
        throw std::runtime_error( "gather_probe: bad arguments" );
    DevBuf<U4> buf;
    DevBuf<unsigned int> sink;
    const unsigned long long nBlocks = (unsigned long long)buffer_bytes / 64;
    buf.reserve( (size_t)nBlocks * 4 );
    sink.reserve( 1 );
    MA_CUDA( cudaMemsetAsync( buf.p, 1, nBlocks * 64, ctx->stream ) );
    const int grid = ctx->num_sms * 8, perThread = 256;
    float best = 1e30f;
    for( int it = 0; it < 4; it++ )
    {
        ctx->timer.start( ctx->stream );
        gather64_kernel<<<grid, 256, 0, ctx->stream>>>( buf.p, nBlocks, perThread, 1234567ull + it, sink.p );
        MA_CUDA( cudaGetLastError( ) );
        const float ms = ctx->timer.stop( ctx->stream );
        ctx->launches++;
        if( it > 0 && ms < best )
            best = ms;
    }
    *gbs = (double)grid * 256 * perThread * 64.0 / ( best * 1e6 );
    MA_API_END
}
