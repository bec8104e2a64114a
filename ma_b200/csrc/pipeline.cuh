// CUDA kernels of the read-alignment pipeline (sm_100a). Each kernel is a thin, persistent-grid wrapper around the
// host+device routines in fmindex.cuh / sochar.cuh / nwglue.cuh; ksw.cuh supplies the banded-DP kernel.
//
//   seed_kernel     one thread per read   BinarySeeding (+ drop-off) and seed enumeration        HBM random 64 B reads
//   locate_kernel   one thread per seed   FMIndex::bwt_sa, strand folding, SoC delta             HBM random 64 B reads
//   socharm_kernel  one thread per read   StripOfConsiderationSeeds + Harmonization              latency / FP64
//   nwplan_kernel   one thread per set    reference window + DP problem enumeration
//   ksw_batch_kernel (ksw.cuh)            one warp per DP problem
//   nwasm_kernel    one thread per set    stitching, scoring, dangling-indel removal
//   alnsort_kernel  one thread per read   final std::sort of the read's alignments
//
// Variable-size outputs are bump-allocated from slabs with one atomicAdd per producer; every kernel keeps counting
// after a slab is full, so the host can grow the slab to the exact size and re-run the stage.
#pragma once
#include "ksw.cuh"
#include "ksw_tiny.cuh"
#include "nwglue.cuh"
#include "mapq.cuh"

namespace ma
{

#define MA_NBINS 24 /* 0..14: ksw_batch_kernel (window class x kind), 15: band too wide, 16..21: ksw_qs_kernel */
#define MA_QS_BIN0 16
#define MA_TINY_BIN 22 /* ksw_tiny_kernel: gap fills whose band covers the whole (<= 32 x 32) matrix */
// control block in device memory; every counter that many threads bump at the same time sits in its own 128-byte
// line (same-line atomics serialise in the L2 slice that owns the line)
struct PipeCtrl
{
    alignas( 128 ) unsigned long long seed_cursor; // seeds allocated
    alignas( 128 ) unsigned long long set_seed_cursor; // harmonized seeds allocated
    alignas( 128 ) unsigned long long set_cursor; // set headers allocated
    alignas( 128 ) unsigned long long task_cursor; // DP tasks allocated
    alignas( 128 ) unsigned long long run_cursor; // alignment run words allocated
    alignas( 128 ) unsigned long long scratch_cursor; // soc/harm scratch bytes
    alignas( 128 ) int next_read;
    alignas( 128 ) int next_read2;
    alignas( 128 ) int next_set;
    alignas( 128 ) int next_set2;
    alignas( 128 ) int next_read3;
    alignas( 128 ) unsigned long long n_ext; // FMIndex::extend_backward calls (roofline unit)
    unsigned long long n_lookup; // ... of which read the occurrence table
    unsigned long long n_invpsi; // bwt_invPsi steps
    unsigned long long n_dropped; // reads cleared by the seeding drop-off heuristic
    int overflow_lists, overflow_fseg, overflow_runs, overflow_pair;
    int max_reported; // largest MappingQuality result vector of the batch
    int n_failed; // reads with a non-zero ReadInfo::status
    unsigned long long reported_cursor; // records of the compact (reported-only) alignment array
    // DP task bins: window class (5) x kind (exact / early-stop left / early-stop right), + 1 for "band too wide"
    alignas( 128 ) int bin_count[ MA_NBINS ];
    alignas( 128 ) unsigned long long bin_tb[ MA_NBINS ];
    alignas( 128 ) int bin_cig[ MA_NBINS ];
};

// read offsets of a batch checked on the device (ma_b200_align_batch: the host does not walk the array again):
// out[0] = longest read, out[1] = 1 if an offset is negative, decreasing, beyond `total`, or a read longer than 2^30 - 1
__global__ void __launch_bounds__( 256 ) offsets_check_kernel( const long long* off, long long n, long long total, int* out )
{
    int maxL = 0, bad = 0;
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x )
    {
        const long long a = off[ i ], b = off[ i + 1 ];
        const long long L = b - a;
        if( a < 0 || L < 0 || L > 0x3fffffff || b > total )
            bad = 1;
        else
            maxL = L > maxL ? (int)L : maxL;
    }
    maxL = __reduce_max_sync( 0xffffffffu, maxL ), bad = __reduce_max_sync( 0xffffffffu, bad );
    if( ( threadIdx.x & 31 ) == 0 )
    {
        if( maxL )
            atomicMax( &out[ 0 ], maxL );
        if( bad )
            atomicExch( &out[ 1 ], 1 );
    }
}

struct ReadInfo // per read
{
    long long seed_off; // first seed of the read in the seed slab
    int n_seeds;
    int set_off; // first set header
    int n_sets;
    int status; // 0, or MA_READ_* bits: the read ran into a capacity of this implementation and has no / partial results
};
#define MA_READ_ELISTS 1 /* more SMEM interval-list entries than the per-read capacity (seeding) */
#define MA_READ_ESEGMENTS 2 /* more filtered segments than the per-read capacity (seeding) */
#define MA_READ_ESETS 4 /* more harmonized seed sets than MA_MAX_SETS_PER_READ */
#define MA_READ_EBAND 8 /* a DP problem of one of its seed sets needs a band wider than the largest window */

struct SetHeader
{
    int read, ordinal;
    unsigned int soc_index;
    int n;
    long long seed_off;
    int task_off, n_tasks;
    unsigned long long win_begin, win_end;
    int valid, pad;
};

struct FSeg // segment that passed the ExtractSeeds filter
{
    int start, size;
    long long sa_start;
    int sa_size, pad;
};

struct SeedKernelArgs
{
    DevIndex I;
    SeedParams P;
    const unsigned char* reads;
    const long long* read_off;
    int n_reads;
    ReadInfo* info;
    DSeed* seeds;
    long long seed_cap;
    SegRec* lists; // per thread: 2 * list_cap
    int list_cap;
    FSeg* fsegs; // per thread: fseg_cap
    int fseg_cap;
    // debug: all segments of every read (tests only)
    SegRec* dbg_segs;
    int* dbg_nsegs;
    int dbg_cap;
    PipeCtrl* ctrl;
    // ma_b200_align_batch: the bytes [0, *reads_ready) of the read slab have arrived (the upload runs under this kernel,
    // in chunks that end on 128-byte lines); nullptr: all resident
    const unsigned long long* reads_ready;
};

struct SeedSink
{
    const SeedParams& P;
    FSeg* fsegs;
    int cap, n = 0;
    long long nSeeds = 0;
    unsigned long long dropSum = 0;
    bool overflow = false;
    SegRec* dbg;
    int dbgCap, nAll = 0;
    __device__ SeedSink( const SeedParams& P, FSeg* f, int cap, SegRec* dbg, int dbgCap )
        : P( P ), fsegs( f ), cap( cap ), dbg( dbg ), dbgCap( dbgCap )
    {}
    __device__ void seg( const SegRec& r )
    {
        if( dbg && nAll < dbgCap )
            dbg[ nAll ] = r;
        nAll++;
        if( P.drop_min_size != 0 )
            dropSum += (unsigned long long)r.size / (unsigned long long)P.drop_min_size; // segment.h:278-289
        // SegmentVector::forEachSeed filter (segment.h:321-336); bSkip is hard-wired to true (:365)
        if( (unsigned long long)r.size < (unsigned long long)P.min_seed_len )
            return;
        if( r.sa.size > (long long)P.max_amb && P.max_amb != 0 )
            return;
        if( n < cap )
            fsegs[ n ] = FSeg{ r.start, r.size, r.sa.start, (int)r.sa.size, 0 };
        else
            overflow = true;
        n++;
        nSeeds += r.sa.size;
    }
};

// Every lane runs the resumable SeederSM of its current read; all lanes of a warp meet at the single
// extend_backward call site, so the warp always has up to 64 independent 64-byte occ-block loads in flight.
// The SMEM interval list of every thread lives in shared memory (first MA_SEED_K entries, 20 bytes each, see SegList);
// only longer lists touch the thread's global scratch.
#ifndef MA_SEED_BLOCK
#define MA_SEED_BLOCK 64
#endif
#ifndef MA_SEED_K
#define MA_SEED_K 12
#endif
#ifndef MA_SEED_MINB
#define MA_SEED_MINB 10 /* 96 registers, 10 CTAs of 64 threads per SM (measured: 25.0 vs 25.7 ms per 1 M reads with 1) */
#endif
__global__ void __launch_bounds__( MA_SEED_BLOCK, MA_SEED_MINB ) seed_kernel( SeedKernelArgs A )
{
    __shared__ U4 sPk[ MA_SEED_K * MA_SEED_BLOCK ];
    __shared__ int sSz[ MA_SEED_K * MA_SEED_BLOCK ];
    __shared__ unsigned short sMu[ MA_SEED_K * MA_SEED_BLOCK ];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    SegRec* la = A.lists + (size_t)tid * ( 2 * A.list_cap + 8 );
    FSeg* fs = A.fsegs + (size_t)tid * A.fseg_cap;
    unsigned long long nExtLocal = 0, nLookupLocal = 0, nDropped = 0;
    SeedSink sink( A.P, fs, A.fseg_cap, nullptr, A.dbg_cap );
    // the first 2 x 40 ints of the thread's list scratch hold the interval stack
    SeederSM<SeedSink> S( A.I, A.P, A.reads, 0,
                          SegList{ sPk + threadIdx.x, sSz + threadIdx.x, sMu + threadIdx.x, MA_SEED_BLOCK, MA_SEED_K, la + 8,
                                   MA_SEED_K + 2 * A.list_cap },
                          sink, (int*)la );
    int read = -1; // -1: needs a new read, -2: queue exhausted
    int L = 0;
    while( true )
    {
        if( read == -1 )
        {
            read = atomicAdd( &A.ctrl->next_read, 1 );
            if( read >= A.n_reads )
                read = -2;
            else
            {
                const long long off = A.read_off[ read ];
                L = (int)( A.read_off[ read + 1 ] - off );
                if( A.reads_ready )
                { // every cache line this read touches must have arrived completely (lines are not re-fetched)
                    const unsigned long long need = ( (unsigned long long)( off + L ) + 127ull ) & ~127ull;
                    while( *( (const volatile unsigned long long*)A.reads_ready ) < need )
                        __nanosleep( 500 );
                }
                sink.n = 0, sink.nSeeds = 0, sink.dropSum = 0, sink.overflow = false, sink.nAll = 0;
                sink.dbg = A.dbg_segs ? A.dbg_segs + (size_t)read * A.dbg_cap : nullptr;
                S.begin( A.reads + off, L );
            }
        }
        SAI rIk;
        int rC = 0;
        bool need = false;
        if( read >= 0 )
        {
            need = S.request( rIk, rC );
            if( !need )
            { // the read is finished: drop-off heuristic (binarySeeding.cpp:172-175) and seed enumeration
                nExtLocal += (unsigned long long)S.nExt, nLookupLocal += (unsigned long long)S.nLookup;
                // a read that exceeds a per-read capacity is reported as such and gets no seeds; the batch goes on
                const int status = ( S.overflow ? MA_READ_ELISTS : 0 ) | ( sink.overflow ? MA_READ_ESEGMENTS : 0 );
                if( status )
                    atomicAdd( &A.ctrl->n_failed, 1 );
                const bool bClear = !A.P.disable_heuristics && A.P.drop_min_size != 0 &&
                                    (double)sink.dropSum < A.P.drop_factor * (double)L &&
                                    (unsigned long long)A.P.genome_size_disable < (unsigned long long)A.I.ref_len;
                if( bClear )
                    nDropped++;
                if( A.dbg_nsegs )
                    A.dbg_nsegs[ read ] = bClear ? 0 : sink.nAll;
                const long long nSeeds = ( bClear || status ) ? 0 : sink.nSeeds;
                long long so = 0;
                if( nSeeds > 0 )
                    so = (long long)atomicAdd( &A.ctrl->seed_cursor, (unsigned long long)nSeeds );
                ReadInfo ri;
                ri.seed_off = so, ri.n_seeds = (int)nSeeds, ri.set_off = 0, ri.n_sets = 0, ri.status = status;
                A.info[ read ] = ri;
                if( nSeeds > 0 && so + nSeeds <= A.seed_cap )
                {
                    long long k = so;
                    const int nf = sink.n;
                    for( int i = 0; i < nf; i++ )
                    {
                        const FSeg f = fs[ i ];
                        for( int j = 0; j < f.sa_size; j++ )
                        {
                            DSeed d;
                            d.q = f.start, d.len = f.size + 1;
                            d.r = f.sa_start + j; // SA row, resolved by locate_kernel
                            d.amb = (unsigned int)f.sa_size, d.fw = 1, d.delta = read; // read id, for locate_kernel
                            A.seeds[ k++ ] = d;
                        }
                    }
                }
                read = -1;
            }
        }
        if( __all_sync( 0xffffffffu, read == -2 ) )
            break;
        __syncwarp( );
        if( need )
            S.consume( extend_backward( A.I, rIk, rC ) );
    }
    if( nExtLocal )
        atomicAdd( &A.ctrl->n_ext, nExtLocal ), atomicAdd( &A.ctrl->n_lookup, nLookupLocal );
    if( nDropped )
        atomicAdd( &A.ctrl->n_dropped, nDropped );
}

struct LocateArgs
{
    DevIndex I;
    DSeed* seeds;
    long long n_seeds;
    const long long* read_off;
    PipeCtrl* ctrl;
};

// Segment::forEachSeed (segment.h:89-113) + ExtractSeeds::setDeltaOfSeed (stripOfConsideration.h:42-54, 97-112)
__global__ void __launch_bounds__( 256 ) locate_kernel( LocateArgs A )
{
    unsigned long long steps = 0;
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < A.n_seeds;
         i += (long long)gridDim.x * blockDim.x )
    {
        DSeed d = A.seeds[ i ];
        const int read = (int)d.delta;
        const long long qlen = A.read_off[ read + 1 ] - A.read_off[ read ];
        int st = 0;
        long long r = bwt_sa( A.I, d.r, &st );
        steps += (unsigned long long)st;
        const bool fw = r < A.I.ref_len / 2;
        if( !fw )
            r = A.I.ref_len - r - 1;
        d.r = r, d.fw = fw ? 1 : 0;
        d.delta = r + ( qlen - d.q ) + ( qlen + 1 ) * seq_id_for_position( A.I, r );
        A.seeds[ i ] = d;
    }
    // warp-aggregate the accounting counter
    for( int o = 16; o > 0; o >>= 1 )
        steps += __shfl_xor_sync( 0xffffffffu, steps, o );
    if( ( threadIdx.x & 31 ) == 0 && steps )
        atomicAdd( &A.ctrl->n_invpsi, steps );
}

struct SocHarmArgs
{
    DevIndex I;
    HarmParams P;
    const long long* read_off;
    int n_reads;
    ReadInfo* info;
    DSeed* seeds;
    unsigned char* scratch; // arena, bump allocated per read
    unsigned long long scratch_cap;
    DSeed* set_seeds;
    long long set_seed_cap;
    SetHeader* sets;
    long long set_cap;
    unsigned int srand_base;
    PipeCtrl* ctrl;
    // the stage as two kernels (socbuild_kernel -> harmonize_kernel): per read, where its scratch starts (~0: none) and
    // how many windows soc_build left in the queue
    unsigned long long* soc_scratch;
    int* soc_nmax;
};

#define MA_MAX_SETS_PER_READ 128

struct DevSetSink
{
    DSeed* slab;
    long long cap;
    PipeCtrl* ctrl;
    int read;
    // per-thread headers of the read under construction
    long long off[ MA_MAX_SETS_PER_READ ];
    int n[ MA_MAX_SETS_PER_READ ];
    unsigned int soc[ MA_MAX_SETS_PER_READ ];
    int count = 0;
    __device__ void set( const DSeed* p, int m, unsigned int socIndex )
    {
        const long long o = (long long)atomicAdd( &ctrl->set_seed_cursor, (unsigned long long)m );
        if( o + m <= cap )
            for( int i = 0; i < m; i++ )
                slab[ o + i ] = p[ i ];
        if( count < MA_MAX_SETS_PER_READ )
            off[ count ] = o, n[ count ] = m, soc[ count ] = socIndex;
        count++; // (the caller checks the final count: pop_back may bring it back under the capacity)
    }
    __device__ void pop_back( unsigned int counter, unsigned int minTries )
    {
        for( unsigned int ui = 0; ui < counter && (unsigned int)count > minTries; ui++ )
            count--;
    }
};

#ifndef MA_SOC_BLOCK
#define MA_SOC_BLOCK 128
#endif
#ifndef MA_SOC_MINB
#define MA_SOC_MINB 12
#endif
__global__ void __launch_bounds__( MA_SOC_BLOCK, MA_SOC_MINB ) socharm_kernel( SocHarmArgs A )
{
    while( true )
    {
        const int read = atomicAdd( &A.ctrl->next_read2, 1 );
        if( read >= A.n_reads )
            break;
        ReadInfo ri = A.info[ read ];
        ri.n_sets = 0, ri.set_off = 0;
        if( ri.n_seeds > 0 )
        {
            const int n = ri.n_seeds;
            // the SoC sorts its seeds in place: work on a copy so that the stage can be re-run after a slab grew
            const size_t copyBytes = ( (size_t)n * sizeof( DSeed ) + 15 ) & ~(size_t)15;
            const size_t need = harm_scratch_need( (size_t)n ) + copyBytes + 64;
            const unsigned long long so = atomicAdd( &A.ctrl->scratch_cursor, (unsigned long long)need );
            if( so + need <= A.scratch_cap )
            {
                DSeed* S = (DSeed*)( A.scratch + so );
                for( int i = 0; i < n; i++ )
                    S[ i ] = A.seeds[ ri.seed_off + i ];
                HarmScratch W = harm_scratch_carve( A.scratch + so + copyBytes, (size_t)n );
                DevSetSink sink;
                sink.slab = A.set_seeds, sink.cap = A.set_seed_cap, sink.ctrl = A.ctrl, sink.read = read;
                const int qlen = (int)( A.read_off[ read + 1 ] - A.read_off[ read ] );
                soc_harm_read( A.I, A.P, S, n, qlen, A.srand_base + (unsigned int)read, W, sink, 0 );
                int ns = sink.count;
                if( ns > MA_MAX_SETS_PER_READ )
                { // reported per read (ma_b200_set_params rejects max_num_soc > MA_MAX_SETS_PER_READ, so: not reachable)
                    if( !( ri.status & MA_READ_ESETS ) )
                        atomicAdd( &A.ctrl->n_failed, 1 );
                    ns = 0, ri.status |= MA_READ_ESETS;
                }
                if( ns > 0 )
                {
                    const long long ho = (long long)atomicAdd( &A.ctrl->set_cursor, (unsigned long long)ns );
                    ri.set_off = (int)ho, ri.n_sets = ns;
                    if( ho + ns <= A.set_cap )
                        for( int i = 0; i < ns; i++ )
                        {
                            SetHeader h;
                            h.read = read, h.ordinal = i, h.soc_index = sink.soc[ i ], h.n = sink.n[ i ];
                            h.seed_off = sink.off[ i ], h.task_off = 0, h.n_tasks = 0;
                            h.win_begin = h.win_end = 0, h.valid = 0, h.pad = 0;
                            A.sets[ ho + i ] = h;
                        }
                }
            }
        }
        A.info[ read ] = ri;
    }
}

// The same stage as two kernels. The one-kernel form is 16 k instructions that every warp walks at its own pace (ncu:
// issue 16 %, instruction-fetch stalls dominant); split, the warps of a launch stay within one half of that code: sort +
// sweep + heap in the first kernel, RANSAC + linesweeps + filters in the second.
__global__ void __launch_bounds__( MA_SOC_BLOCK, MA_SOC_MINB ) socbuild_kernel( SocHarmArgs A )
{
    while( true )
    {
        const int read = atomicAdd( &A.ctrl->next_read2, 1 );
        if( read >= A.n_reads )
            break;
        const ReadInfo ri = A.info[ read ];
        unsigned long long where = ~0ull;
        int nMax = 0;
        if( ri.n_seeds > 0 )
        {
            const int n = ri.n_seeds;
            // the SoC sorts its seeds in place: work on a copy so that the stage can be re-run after a slab grew
            const size_t copyBytes = ( (size_t)n * sizeof( DSeed ) + 15 ) & ~(size_t)15;
            const size_t need = harm_scratch_need( (size_t)n ) + copyBytes + 64;
            const unsigned long long so = atomicAdd( &A.ctrl->scratch_cursor, (unsigned long long)need );
            if( so + need <= A.scratch_cap )
            {
                DSeed* S = (DSeed*)( A.scratch + so );
                for( int i = 0; i < n; i++ )
                    S[ i ] = A.seeds[ ri.seed_off + i ];
                HarmScratch W = harm_scratch_carve( A.scratch + so + copyBytes, (size_t)n );
                const int qlen = (int)( A.read_off[ read + 1 ] - A.read_off[ read ] );
                nMax = soc_build( A.I, A.P, S, n, qlen, W.maxima, W.vref );
                where = so;
            }
        }
        A.soc_scratch[ read ] = where, A.soc_nmax[ read ] = nMax;
    }
}

__global__ void __launch_bounds__( MA_SOC_BLOCK, MA_SOC_MINB ) harmonize_kernel( SocHarmArgs A )
{
    while( true )
    {
        const int read = atomicAdd( &A.ctrl->next_read3, 1 );
        if( read >= A.n_reads )
            break;
        ReadInfo ri = A.info[ read ];
        ri.n_sets = 0, ri.set_off = 0;
        const unsigned long long so = A.soc_scratch[ read ];
        if( ri.n_seeds > 0 && so != ~0ull )
        {
            const int n = ri.n_seeds;
            const size_t copyBytes = ( (size_t)n * sizeof( DSeed ) + 15 ) & ~(size_t)15;
            const DSeed* S = (const DSeed*)( A.scratch + so );
            HarmScratch W = harm_scratch_carve( A.scratch + so + copyBytes, (size_t)n );
            DevSetSink sink;
            sink.slab = A.set_seeds, sink.cap = A.set_seed_cap, sink.ctrl = A.ctrl, sink.read = read;
            const int qlen = (int)( A.read_off[ read + 1 ] - A.read_off[ read ] );
            soc_harm_pops( A.I, A.P, S, n, qlen, A.srand_base + (unsigned int)read, W, sink, A.soc_nmax[ read ] );
            int ns = sink.count;
            if( ns > MA_MAX_SETS_PER_READ )
            { // reported per read (ma_b200_set_params rejects max_num_soc > MA_MAX_SETS_PER_READ, so: not reachable)
                if( !( ri.status & MA_READ_ESETS ) )
                    atomicAdd( &A.ctrl->n_failed, 1 );
                ns = 0, ri.status |= MA_READ_ESETS;
            }
            if( ns > 0 )
            {
                const long long ho = (long long)atomicAdd( &A.ctrl->set_cursor, (unsigned long long)ns );
                ri.set_off = (int)ho, ri.n_sets = ns;
                if( ho + ns <= A.set_cap )
                    for( int i = 0; i < ns; i++ )
                    {
                        SetHeader h;
                        h.read = read, h.ordinal = i, h.soc_index = sink.soc[ i ], h.n = sink.n[ i ];
                        h.seed_off = sink.off[ i ], h.task_off = 0, h.n_tasks = 0;
                        h.win_begin = h.win_end = 0, h.valid = 0, h.pad = 0;
                        A.sets[ ho + i ] = h;
                    }
            }
        }
        A.info[ read ] = ri;
    }
}

struct NwPlanArgs
{
    DevIndex I;
    NwParams P;
    const long long* read_off;
    SetHeader* sets;
    int n_sets;
    const DSeed* set_seeds;
    KswTask* tasks;
    long long task_cap;
    int* bin_order; // [n_bins][task_cap] task ids per window bin
    PipeCtrl* ctrl;
    ReadInfo* info;
};

// One thread per seed set: window, task count, one warp-aggregated slot allocation, tasks.
__global__ void __launch_bounds__( 128 ) nwplan_kernel( NwPlanArgs A )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane of a warp takes part in the aggregated allocation of each round
    for( int base = ( blockIdx.x * blockDim.x + threadIdx.x ) & ~31; base < A.n_sets; base += gridDim.x * blockDim.x )
    {
        const int si = base + lane;
        const bool live = si < A.n_sets;
        SetHeader h;
        const DSeed* S = nullptr;
        long long qbase = 0;
        int qlen = 0, nTasks = 0;
        NwWindow w;
        w.valid = false;
        if( live )
        {
            h = A.sets[ si ];
            S = A.set_seeds + h.seed_off;
            qbase = A.read_off[ h.read ];
            qlen = (int)( A.read_off[ h.read + 1 ] - qbase );
            w = nw_window( A.I, A.P, S, h.n );
            h.valid = w.valid ? 1 : 0, h.win_begin = w.beginRef, h.win_end = w.endRef, h.n_tasks = 0, h.task_off = 0;
            if( w.valid )
            {
                NwPlanner cnt( A.P, nullptr, qbase, (long long)w.beginRef );
                nw_walk( S, h.n, qlen, w, cnt );
                nTasks = cnt.n;
            }
        }
        // exclusive scan of the task counts over the warp, one atomic for all of them
        int incl = nTasks;
        for( int o = 1; o < 32; o <<= 1 )
        {
            const int v = __shfl_up_sync( FULL, incl, o );
            if( lane >= o )
                incl += v;
        }
        const int total = __shfl_sync( FULL, incl, 31 );
        long long wbase = 0;
        if( lane == 0 && total > 0 )
            wbase = (long long)atomicAdd( &A.ctrl->task_cursor, (unsigned long long)total );
        wbase = __shfl_sync( FULL, wbase, 0 );
        if( live )
        {
            if( nTasks > 0 )
            {
                const long long to = wbase + incl - nTasks;
                h.n_tasks = nTasks, h.task_off = (int)to;
                if( to + nTasks <= A.task_cap )
                {
                    NwPlanner pl( A.P, A.tasks + to, qbase, (long long)w.beginRef );
                    nw_walk( S, h.n, qlen, w, pl );
                    // a band wider than the largest window of ksw_batch_kernel: the set yields no alignment and the
                    // read is reported (ReadInfo::status), the batch goes on
                    bool tooWide = false;
                    for( int t = 0; t < nTasks; t++ )
                        tooWide |= ksw_bin_of( ksw_ncol16( A.tasks[ to + t ].qlen, A.tasks[ to + t ].tlen, A.tasks[ to + t ].w ) ) >= 5;
                    if( tooWide )
                    {
                        for( int t = 0; t < nTasks; t++ )
                            A.tasks[ to + t ].tag |= MA_TASK_SKIP;
                        h.valid = 0;
                        if( !( atomicOr( &A.info[ h.read ].status, MA_READ_EBAND ) & MA_READ_EBAND ) )
                            atomicAdd( &A.ctrl->n_failed, 1 );
                    }
                }
            }
            A.sets[ si ] = h;
        }
    }
}

struct NwBinArgs
{
    const KswTask* tasks;
    int n_tasks;
    long long task_cap;
    int* bin_order; // [n_bins][task_cap] task ids per window bin
    PipeCtrl* ctrl;
    KswScore score;
    int use_qs; // route early-stop extensions with short queries to ksw_qs_kernel
    int use_tiny; // route small gap fills to ksw_tiny_kernel
};

// One thread per DP task: window bin, slot in the bin's order list (one atomic per warp and bin), slab sizes of the bin
__global__ void __launch_bounds__( 256 ) nwbin_kernel( NwBinArgs A )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    for( int base = ( blockIdx.x * blockDim.x + threadIdx.x ) & ~31; base < A.n_tasks; base += gridDim.x * blockDim.x )
    {
        const int ti = base + lane;
        int b = MA_NBINS; // lanes beyond the end form their own group
        unsigned int tb = 0, cig = 0;
        if( ti < A.n_tasks )
        {
            const KswTask T = A.tasks[ ti ];
            const int nc = ksw_ncol16( T.qlen, T.tlen, T.w );
            const bool skip = ( T.tag & MA_TASK_SKIP ) != 0;
            // all warps of a launch should run the same instantiation of the row loop (the kernel is large and a mix
            // of code paths thrashes the instruction cache): tasks are binned by window class AND kind
            const int wc = ksw_bin_of( nc );
            const int kind = !( T.tag & MA_TASK_EARLYSTOP ) ? 0 : ( T.flag & MA_KSW_RIGHT ) ? 2 : 1;
            b = wc < 5 ? wc * 3 + kind : 15;
            unsigned long long bytes = ( ( (unsigned long long)T.qlen + T.tlen ) * nc + 255 ) >> 8;
            const int qs = A.use_qs ? ksw_qs_class( A.score, T.qlen, T.tlen, T.w, T.tag ) : 0;
            if( qs > 0 )
            {
                b = MA_QS_BIN0 + ( qs - 1 ) * 2 + ( kind == 2 ? 1 : 0 );
                bytes = (unsigned long long)ksw_qs_tb_bytes( qs, T.qlen, T.tlen, T.w ) >> 8;
            }
            tb = bytes > 0xffffffffull ? 0xffffffffu : (unsigned int)bytes; // in units of 256 bytes
            cig = (unsigned int)( ( T.qlen + T.tlen + 2 + 63 ) & ~63 );
            if( A.use_tiny && ksw_tiny_ok( A.score, T.qlen, T.tlen, T.w, T.flag, T.tag ) )
                b = MA_TINY_BIN, tb = 0, cig = 0; // one thread per problem, state in thread-local memory
            if( skip )
                b = MA_NBINS, tb = 0, cig = 0;
        }
        const unsigned m = __match_any_sync( FULL, b );
        const int leader = __ffs( m ) - 1;
        const unsigned mtb = __reduce_max_sync( m, tb ), mcig = __reduce_max_sync( m, cig );
        int slot0 = 0;
        if( lane == leader && b < MA_NBINS )
        {
            slot0 = atomicAdd( &A.ctrl->bin_count[ b ], __popc( m ) );
            atomicMax( &A.ctrl->bin_tb[ b ], (unsigned long long)mtb << 8 );
            atomicMax( &A.ctrl->bin_cig[ b ], (int)mcig );
        }
        slot0 = __shfl_sync( FULL, slot0, leader );
        if( b < MA_NBINS )
            A.bin_order[ (long long)b * A.task_cap + slot0 + __popc( m & ( ( 1u << lane ) - 1 ) ) ] = ti;
    }
}

struct NwAsmArgs
{
    DevIndex I;
    NwParams P;
    const unsigned char* reads;
    const long long* read_off;
    const SetHeader* sets;
    int n_sets;
    const DSeed* set_seeds;
    const KswOut* res;
    const unsigned int* cigar;
    DAln* alns; // one per set, same index
    unsigned int* runs;
    long long run_cap;
    unsigned int* run_scratch; // per thread
    int run_scratch_cap;
    PipeCtrl* ctrl;
};

__global__ void __launch_bounds__( 128 ) nwasm_kernel( NwAsmArgs A )
{
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int* scratch = A.run_scratch + (size_t)tid * A.run_scratch_cap;
    while( true )
    {
        const int si = atomicAdd( &A.ctrl->next_set, 1 );
        if( si >= A.n_sets )
            break;
        const SetHeader h = A.sets[ si ];
        DAln a;
        a.begin_ref = a.end_ref = a.score = 0, a.begin_q = a.end_q = a.length = a.n_runs = 0;
        a.soc_index = h.soc_index, a.read = h.read, a.run_off = 0, a.rank = h.ordinal, a.flags = 0;
        a.mapq = NAN, a.rank_mq = -1, a.pair_rank = -1;
        if( h.valid )
        {
            const long long qbase = A.read_off[ h.read ];
            const int qlen = (int)( A.read_off[ h.read + 1 ] - qbase );
            const DSeed* S = A.set_seeds + h.seed_off;
            NwWindow w{ h.win_begin, h.win_end, true };
            NwAssembler as( A.I, A.P, A.reads + qbase, h.win_begin, A.res + h.task_off, A.cigar, scratch,
                            A.run_scratch_cap );
            nw_walk( S, h.n, qlen, w, as );
            as.removeDangeling( );
            if( as.overflow )
                atomicExch( &A.ctrl->overflow_runs, 1 );
            a.begin_ref = (long long)as.beginR, a.end_ref = (long long)as.endR, a.score = as.score;
            a.begin_q = (int)as.beginQ, a.end_q = (int)as.endQ, a.length = (int)as.length;
            a.n_runs = as.nRuns - as.front;
            if( a.n_runs > 0 )
            {
                const long long ro = (long long)atomicAdd( &A.ctrl->run_cursor, (unsigned long long)a.n_runs );
                a.run_off = ro;
                if( ro + a.n_runs <= A.run_cap )
                    for( int i = 0; i < a.n_runs; i++ )
                        A.runs[ ro + i ] = scratch[ as.front + i ];
            }
        }
        A.alns[ si ] = a;
    }
}

struct AlnSortArgs
{
    const ReadInfo* info;
    int n_reads;
    DAln* alns;
    PipeCtrl* ctrl;
};

// the final std::sort of NeedlemanWunsch::execute with Alignment::larger (needlemanWunsch.h:131-132)
__global__ void __launch_bounds__( 128 ) alnsort_kernel( AlnSortArgs A )
{
    for( int read = blockIdx.x * blockDim.x + threadIdx.x; read < A.n_reads; read += gridDim.x * blockDim.x )
    {
        const ReadInfo ri = A.info[ read ];
        if( ri.n_sets <= 0 )
            continue;
        int ord[ MA_MAX_SETS_PER_READ ];
        const int n = ri.n_sets;
        DAln* al = A.alns + ri.set_off;
        for( int i = 0; i < n; i++ )
            ord[ i ] = i;
        stl::sort( ord, ord + n, [ & ]( int a, int b ) {
            if( al[ a ].score == al[ b ].score )
                return al[ a ].soc_index < al[ b ].soc_index;
            return al[ a ].score > al[ b ].score;
        } );
        for( int i = 0; i < n; i++ )
            al[ ord[ i ] ].rank = i;
    }
}


struct MapqArgs
{
    MapqParams P;
    const ReadInfo* info;
    const long long* read_off;
    int n_reads;
    DAln* alns;
    const unsigned int* runs;
    long long ref_len;
    long long* pair_sc; // per thread: pair_cap candidate scores
    int* pair_meta; // per thread: 2 * pair_cap ints
    int pair_cap;
    PipeCtrl* ctrl;
};

// Reported-only output (ma_b200_set_reported_only): the records MappingQuality / PairedReads hand to the writer
// (rank_mq >= 0) of every read, copied next to each other into a second array, with the read's info pointing at them. On a
// human-sized genome a read has ~3 seed sets but ~1.1 reported alignments: the download of the records shrinks 3x.
struct CompactArgs
{
    const ReadInfo* info;
    ReadInfo* info_out;
    int n_reads;
    const DAln* alns;
    DAln* alns_out;
    PipeCtrl* ctrl;
};
__global__ void __launch_bounds__( 128 ) compact_reported_kernel( CompactArgs A )
{
    for( int read = blockIdx.x * blockDim.x + threadIdx.x; read < A.n_reads; read += gridDim.x * blockDim.x )
    {
        ReadInfo ri = A.info[ read ];
        int n = 0;
        for( int k = 0; k < ri.n_sets; k++ )
            n += A.alns[ ri.set_off + k ].rank_mq >= 0 ? 1 : 0;
        long long o = 0;
        if( n > 0 )
            o = (long long)atomicAdd( &A.ctrl->reported_cursor, (unsigned long long)n );
        int w = 0;
        for( int k = 0; k < ri.n_sets && w < n; k++ )
            if( A.alns[ ri.set_off + k ].rank_mq >= 0 )
                A.alns_out[ o + w++ ] = A.alns[ ri.set_off + k ];
        ri.set_off = (int)o, ri.n_sets = n;
        A.info_out[ read ] = ri;
    }
}

// MappingQuality::execute, one thread per read; also records the largest result vector (sizes the pairing scratch)
__global__ void __launch_bounds__( 128 ) mapq_kernel( MapqArgs A )
{
    int ord[ MA_MAX_SETS_PER_READ ];
    int most = 0;
    for( int read = blockIdx.x * blockDim.x + threadIdx.x; read < A.n_reads; read += gridDim.x * blockDim.x )
    {
        const ReadInfo ri = A.info[ read ];
        const int m = mapping_quality_read( A.P, A.alns + ri.set_off, ri.n_sets, A.runs,
                                            A.read_off[ read + 1 ] - A.read_off[ read ], ord );
        most = m > most ? m : most;
    }
    most = __reduce_max_sync( 0xffffffffu, most );
    if( ( threadIdx.x & 31 ) == 0 && most > 0 )
        atomicMax( &A.ctrl->max_reported, most );
}

// PairedReads::execute, one thread per pair of consecutive reads (2k, 2k + 1); a trailing single read keeps its vector
__global__ void __launch_bounds__( 128 ) pair_kernel( MapqArgs A )
{
    int ord1[ MA_MAX_SETS_PER_READ ], ord2[ MA_MAX_SETS_PER_READ ];
    const int tid = blockIdx.x * blockDim.x + threadIdx.x;
    long long* sc = A.pair_sc + (size_t)tid * A.pair_cap;
    int* meta = A.pair_meta + (size_t)tid * 2 * A.pair_cap;
    const int nPairs = ( A.n_reads + 1 ) / 2;
    for( int p = tid; p < nPairs; p += gridDim.x * blockDim.x )
    {
        const int r1 = 2 * p, r2 = 2 * p + 1;
        const ReadInfo i1 = A.info[ r1 ];
        ReadInfo i2;
        i2.set_off = 0, i2.n_sets = 0;
        long long q2 = 0;
        if( r2 < A.n_reads )
            i2 = A.info[ r2 ], q2 = A.read_off[ r2 + 1 ] - A.read_off[ r2 ];
        const int rc = paired_reads_pair( A.P, A.ref_len, A.alns + i1.set_off, i1.n_sets,
                                          A.read_off[ r1 + 1 ] - A.read_off[ r1 ], A.alns + i2.set_off, i2.n_sets, q2,
                                          A.runs, ord1, ord2, sc, meta, A.pair_cap );
        if( rc < 0 )
            atomicExch( &A.ctrl->overflow_pair, 1 );
    }
}

// Roofline probe for the seeding kernels: independent random 64-byte block reads over a buffer of
// the index' size, the same access shape as bwt_occ4 (four 128-bit loads of one 64-byte line), no dependent chain.
__global__ void __launch_bounds__( 256 ) gather64_kernel( const U4* buf, unsigned long long nBlocks, int perThread,
                                                         unsigned long long seed, unsigned int* sink )
{
    unsigned long long x = seed + 0x9E3779B97F4A7C15ull * ( (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x + 1 );
    unsigned int acc = 0;
    for( int i = 0; i < perThread; i += 4 )
    {
        U4 v[ 4 ][ 4 ];
#pragma unroll
        for( int j = 0; j < 4; j++ )
        { // 4 independent blocks in flight per thread
            x ^= x >> 12, x ^= x << 25, x ^= x >> 27;
            const unsigned long long b = ( ( x * 0x2545F4914F6CDD1Dull ) >> 11 ) % nBlocks;
            const U4* p = buf + 4 * b;
            v[ j ][ 0 ] = ld_u4( p ), v[ j ][ 1 ] = ld_u4( p + 1 ), v[ j ][ 2 ] = ld_u4( p + 2 ), v[ j ][ 3 ] = ld_u4( p + 3 );
        }
#pragma unroll
        for( int j = 0; j < 4; j++ )
            acc += v[ j ][ 0 ].x ^ v[ j ][ 1 ].y ^ v[ j ][ 2 ].z ^ v[ j ][ 3 ].w;
    }
    if( acc == 0x12345678u )
        *sink = acc;
}

} // namespace ma
