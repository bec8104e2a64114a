// libma_b200.so — C ABI (include/ma_b200.h) over the sm_100a kernels. Host side of the drop-in boundary.
#include "common.cuh"
#include "ksw.cuh"
#include <algorithm>
#include <numeric>
#include <stdexcept>

using namespace ma;

struct KswHostBin
{
    int W;
    std::vector<int> order;
    long long tb_stride = 0;
    int cig_stride = 0;
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
    DevBuf<unsigned long long> ksw_ctrl; // [0] cigar cursor, [1] next (as int), [2] error (as int)
    std::vector<KswHostBin> ksw_bins;
    unsigned long long ksw_cigar_used = 0;
};

static KswScore make_score( const ma_b200_params& p )
{
    KswScore s;
    s.match = p.match;
    s.mismatch = -p.mismatch;
    int q = p.gap, e = p.extend, q2 = p.gap2, e2 = p.extend2;
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
    // ParameterSetManager(), parameter.h:1079-1104
    if( s == "default" )
        return MA_B200_OK;
    if( s == "illumina" || s == "illuminapaired" )
    {
        p->seeding_technique = 1, p->max_ambiguity = 500, p->min_num_soc = 10, p->max_num_soc = 20;
        return MA_B200_OK;
    }
    if( s == "pacbio" )
    {
        p->min_num_soc = 5;
        return MA_B200_OK;
    }
    if( s == "nanopore" )
    {
        p->seeding_technique = 1, p->min_num_soc = 5;
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

extern "C" void ma_b200_destroy( ma_b200_ctx* ctx )
{
    if( !ctx )
        return;
    cudaSetDevice( ctx->device );
    if( ctx->stream )
        cudaStreamDestroy( ctx->stream );
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

extern "C" int ma_b200_set_params( ma_b200_ctx* ctx, const ma_b200_params* params )
{
    if( !ctx || !params )
        return MA_B200_EINVAL;
    ctx->params = *params;
    return MA_B200_OK;
}

// ------------------------------------------------------------------------------------------------ DP batch
static const int kKswWindows[] = { 128, 256, 512, 1024, 2048 };

template <int W> static long long ksw_bin_grid( ma_b200_ctx* ctx, const KswHostBin& bin, long long tbBudget )
{
    const int warpsPerCta = 8;
    const size_t smem = sizeof( KswSmem<W> ) * warpsPerCta;
    MA_CUDA( cudaFuncSetAttribute( ksw_batch_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem ) );
    int perSm = 0;
    MA_CUDA( cudaOccupancyMaxActiveBlocksPerMultiprocessor( &perSm, ksw_batch_kernel<W>, 256, smem ) );
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
    const size_t smem = sizeof( KswSmem<W> ) * 8;
    ksw_batch_kernel<W><<<(unsigned)grid, 256, smem, ctx->stream>>>( A );
    MA_CUDA( cudaGetLastError( ) );
    ctx->launches++;
}

static void ksw_plan( ma_b200_ctx* ctx, int64_t n, const ma_b200_ksw_task* tasks )
{
    ctx->ksw_bins.clear( );
    for( int W : kKswWindows )
        ctx->ksw_bins.push_back( KswHostBin{ W, { }, 0, 0 } );
    long long bound = 0;
    std::vector<long long> cost( n );
    for( int64_t i = 0; i < n; i++ )
    {
        const auto& t = tasks[ i ];
        if( t.qlen < 0 || t.tlen < 0 )
            throw std::runtime_error( "ksw task with negative length" );
        const int nc = ksw_ncol16( t.qlen, t.tlen, t.w );
        int b = -1;
        for( size_t k = 0; k < ctx->ksw_bins.size( ); k++ )
            if( ctx->ksw_bins[ k ].W >= nc + 48 )
            {
                b = (int)k;
                break;
            }
        if( b < 0 )
            throw std::runtime_error( "ksw task: band wider than the largest supported window (2000 columns)" );
        auto& bin = ctx->ksw_bins[ b ];
        bin.order.push_back( (int)i );
        const long long rows = (long long)t.qlen + t.tlen;
        bin.tb_stride = std::max( bin.tb_stride, ( rows * nc + 255 ) & ~255ll );
        bin.cig_stride = std::max<int>( bin.cig_stride, (int)( ( rows + 2 + 63 ) & ~63ll ) );
        cost[ i ] = rows * nc;
        bound += rows + 2;
    }
    for( auto& bin : ctx->ksw_bins ) // longest first: the dynamic queue then load-balances the tail
        std::stable_sort( bin.order.begin( ), bin.order.end( ),
                          [ & ]( int a, int b ) { return cost[ a ] > cost[ b ]; } );
    ctx->ksw_cigar_bound = bound;
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
    ksw_plan( ctx, n, tasks );
    ctx->ksw_n = n;
    ctx->ksw_tasks.reserve( (size_t)n + 1 );
    ctx->ksw_seq.reserve( (size_t)seq_bytes + 1 );
    ctx->ksw_out.reserve( (size_t)n + 1 );
    ctx->ksw_order.reserve( (size_t)n + 1 );
    ctx->ksw_ctrl.reserve( 4 );
    if( n > 0 )
    {
        MA_CUDA( cudaMemcpyAsync( ctx->ksw_tasks.p, tasks, n * sizeof( KswTask ), cudaMemcpyHostToDevice,
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
    }
    // cigar slab: start with a typical size; ksw_run re-runs with the exact bound if it overflows
    ctx->ksw_cigar_cap = std::min<long long>( ctx->ksw_cigar_bound, std::max<long long>( 48 * n, 1 << 16 ) );
    ctx->ksw_cigar.reserve( (size_t)ctx->ksw_cigar_cap + 1 );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    MA_API_END
}

static int ksw_run_once( ma_b200_ctx* ctx )
{
    MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p, 0, 4 * sizeof( unsigned long long ), ctx->stream ) );
    const KswScore score = make_score( ctx->params );
    const long long budget = 12ll << 30;
    // size the per-warp scratch for the largest bin first: DevBuf::reserve may free + reallocate
    std::vector<long long> grids;
    size_t tbNeed = 0, csNeed = 0;
    for( auto& bin : ctx->ksw_bins )
    {
        long long g = 0;
        if( !bin.order.empty( ) )
            switch( bin.W )
            {
                case 128: g = ksw_bin_grid<128>( ctx, bin, budget ); break;
                case 256: g = ksw_bin_grid<256>( ctx, bin, budget ); break;
                case 512: g = ksw_bin_grid<512>( ctx, bin, budget ); break;
                case 1024: g = ksw_bin_grid<1024>( ctx, bin, budget ); break;
                default: g = ksw_bin_grid<2048>( ctx, bin, budget ); break;
            }
        grids.push_back( g );
        tbNeed = std::max<size_t>( tbNeed, (size_t)( g * 8 * bin.tb_stride ) );
        csNeed = std::max<size_t>( csNeed, (size_t)( g * 8 * bin.cig_stride ) );
    }
    ctx->ksw_tb.reserve( tbNeed + 256 );
    ctx->ksw_cigscratch.reserve( csNeed + 64 );
    size_t o = 0, b = 0;
    for( auto& bin : ctx->ksw_bins )
    {
        const long long grid = grids[ b++ ];
        if( bin.order.empty( ) )
            continue;
        KswBatchArgs A;
        A.tasks = ctx->ksw_tasks.p;
        A.order = ctx->ksw_order.p + o;
        A.n = (int)bin.order.size( );
        A.seq = ctx->ksw_seq.p;
        A.pac = nullptr, A.fwd_len = 0;
        A.out = ctx->ksw_out.p;
        A.cigar = ctx->ksw_cigar.p;
        A.cigar_cap = ctx->ksw_cigar_cap;
        A.cigar_cursor = ctx->ksw_ctrl.p;
        A.tb = ctx->ksw_tb.p;
        A.tb_stride = bin.tb_stride;
        A.cigscratch = ctx->ksw_cigscratch.p;
        A.cigscratch_stride = bin.cig_stride;
        A.next = (int*)( ctx->ksw_ctrl.p + 1 );
        A.error = (int*)( ctx->ksw_ctrl.p + 2 );
        A.score = score;
        MA_CUDA( cudaMemsetAsync( ctx->ksw_ctrl.p + 1, 0, sizeof( unsigned long long ), ctx->stream ) );
        switch( bin.W )
        {
            case 128: launch_ksw_bin<128>( ctx, A, grid ); break;
            case 256: launch_ksw_bin<256>( ctx, A, grid ); break;
            case 512: launch_ksw_bin<512>( ctx, A, grid ); break;
            case 1024: launch_ksw_bin<1024>( ctx, A, grid ); break;
            default: launch_ksw_bin<2048>( ctx, A, grid ); break;
        }
        o += bin.order.size( );
    }
    unsigned long long ctrl[ 3 ];
    MA_CUDA( cudaMemcpyAsync( ctrl, ctx->ksw_ctrl.p, sizeof( ctrl ), cudaMemcpyDeviceToHost, ctx->stream ) );
    MA_CUDA( cudaStreamSynchronize( ctx->stream ) );
    ctx->ksw_cigar_used = ctrl[ 0 ];
    return (int)ctrl[ 2 ];
}

extern "C" int ma_b200_ksw_run( ma_b200_ctx* ctx, float* kernel_ms )
{
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
