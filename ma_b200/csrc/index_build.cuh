// FM-index construction on the GPU (off the measured path; makes the 100 Mbp benchmark index in seconds instead
// of the reference's single-threaded minutes).  Produces bit-identical arrays to FMIndex(pPack):
//   build_FMIndex                libs/ma/src/container/fMIndex.cpp:316-391
//   bwt_bwtupdate_core_step2     fMIndex.cpp:204-264  (64-byte occurrence blocks + trailing counter block)
//   bwt_cal_sa_step3             fMIndex.cpp:266-314  (one SA sample per 32 rows, sa[0] = -1)
//   Pack::vSetNucleotideOnPos    libs/ma/inc/ma/container/pack.h:162-167
// Suffix array of T = forward ++ reverse-complement by prefix doubling with cub radix sorts (the BWT of a text is
// unique, so any correct suffix sort reproduces the reference's bwtLarge / is_bwt output).
#pragma once
#include "common.cuh"
#include "fmindex.cuh"
#include <cub/cub.cuh>

namespace ma
{

__global__ void ib_text_kernel( const unsigned char* fwd, long long n, unsigned char* T )
{
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n;
         i += (long long)gridDim.x * blockDim.x )
        T[ i ] = i < n ? fwd[ i ] : (unsigned char)( 3 - fwd[ 2 * n - 1 - i ] );
}

__global__ void ib_pack_kernel( const unsigned char* fwd, long long n, unsigned char* pac, long long nBytes )
{
    for( long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nBytes;
         b += (long long)gridDim.x * blockDim.x )
    {
        unsigned int v = 0;
        for( int j = 0; j < 4; j++ )
        {
            const long long p = 4 * b + j;
            v |= ( p < n ? ( fwd[ p ] & 3u ) : 0u ) << ( 6 - 2 * j );
        }
        pac[ b ] = (unsigned char)v;
    }
}

#define MA_IB_K0 29 /* bases in the first sort key: 58 bits + 5 bits of valid-length */

__global__ void ib_initkeys_kernel( const unsigned char* T, long long N, unsigned long long* key, unsigned int* idx )
{
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x )
    {
        unsigned long long k = 0;
        const long long valid = N - i < MA_IB_K0 ? N - i : MA_IB_K0;
        for( int j = 0; j < MA_IB_K0; j++ )
            k = ( k << 2 ) | ( j < valid ? (unsigned long long)T[ i + j ] : 0ull );
        key[ i ] = ( k << 5 ) | (unsigned long long)valid; // shorter suffix ($ < A) sorts first among equal prefixes
        idx[ i ] = (unsigned int)i;
    }
}

__global__ void ib_flags_kernel( const unsigned long long* key, long long N, unsigned int* flag )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N;
         j += (long long)gridDim.x * blockDim.x )
        flag[ j ] = ( j == 0 || key[ j ] != key[ j - 1 ] ) ? 1u : 0u;
}

__global__ void ib_scatter_rank_kernel( const unsigned int* idx, const unsigned int* rankSorted, long long N,
                                        unsigned int* rank )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N;
         j += (long long)gridDim.x * blockDim.x )
        rank[ idx[ j ] ] = rankSorted[ j ];
}

__global__ void ib_doublekeys_kernel( const unsigned int* rank, long long N, long long h, unsigned long long* key,
                                      unsigned int* idx )
{
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x )
    {
        const unsigned long long hi = rank[ i ], lo = i + h < N ? rank[ i + h ] : 0u;
        key[ i ] = ( hi << 32 ) | lo;
        idx[ i ] = (unsigned int)i;
    }
}

__global__ void ib_primary_kernel( const unsigned int* sa, long long N, long long* primary )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < N;
         j += (long long)gridDim.x * blockDim.x )
        if( sa[ j ] == 0 )
            *primary = j + 1; // rows are 0..N, row 0 is the $ suffix
}

// one thread per 16-symbol data word of the $-removed BWT
__global__ void ib_bwtwords_kernel( const unsigned char* T, const unsigned int* sa, long long N, long long primary,
                                    unsigned int* dw, long long nw )
{
    for( long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nw;
         w += (long long)gridDim.x * blockDim.x )
    {
        unsigned int v = 0;
        for( int j = 0; j < 16; j++ )
        {
            const long long b = 16 * w + j;
            unsigned int c = 0;
            if( b < N )
            {
                const long long row = b < primary ? b : b + 1;
                const long long p = row == 0 ? N : (long long)sa[ row - 1 ];
                c = T[ p - 1 ]; // p != 0 because the primary row was skipped
            }
            v |= c << ( ( 15 - j ) * 2 );
        }
        dw[ w ] = v;
    }
}

// per 128-symbol block: counts of A,C,G,T among its valid symbols (cnt[c * nblk + b])
__global__ void ib_blockcounts_kernel( const unsigned int* dw, long long nw, long long N, long long nblk,
                                       long long* cnt )
{
    for( long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < nblk;
         b += (long long)gridDim.x * blockDim.x )
    {
        int c = 0, g = 0, t = 0;
        const long long valid = N - 128 * b < 128 ? N - 128 * b : 128;
        for( int j = 0; j < 8; j++ )
            if( 8 * b + j < nw )
                count_word( dw[ 8 * b + j ], (int)( valid - 16 * j ), c, g, t );
        cnt[ 0 * nblk + b ] = valid - c - g - t;
        cnt[ 1 * nblk + b ] = c;
        cnt[ 2 * nblk + b ] = g;
        cnt[ 3 * nblk + b ] = t;
    }
}

// final layout: block b = 4 x u64 exclusive counts + up to 8 data words; one trailing counter block with the totals
__global__ void ib_layout_kernel( const unsigned int* dw, long long nw, const long long* excl, const long long* cnt,
                                  long long nblk, unsigned int* out )
{
    for( long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b <= nblk;
         b += (long long)gridDim.x * blockDim.x )
    {
        unsigned int* o = out + 16 * b;
        if( b == nblk )
            o = out + 8 * nblk + nw; // trailing block follows the last (possibly short) data block
        for( int c = 0; c < 4; c++ )
        {
            const unsigned long long v =
                b < nblk ? (unsigned long long)excl[ c * nblk + b ]
                         : (unsigned long long)( excl[ c * nblk + nblk - 1 ] + cnt[ c * nblk + nblk - 1 ] );
            o[ 2 * c ] = (unsigned int)v, o[ 2 * c + 1 ] = (unsigned int)( v >> 32 );
        }
        if( b < nblk )
            for( int j = 0; j < 8; j++ )
                if( 8 * b + j < nw )
                    o[ 8 + j ] = dw[ 8 * b + j ];
    }
}

__global__ void ib_sasample_kernel( const unsigned int* sa, long long N, int intv, long long* out, long long nOut )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nOut;
         j += (long long)gridDim.x * blockDim.x )
    {
        const long long row = j * intv;
        out[ j ] = row == 0 ? -1 : (long long)sa[ row - 1 ];
    }
}

struct IndexBuildResult
{
    long long primary;
    long long L2[ 5 ];
    long long n_words, n_sa, n_pac;
    int rounds;
};

// fills the context's index buffers; returns primary / L2
inline IndexBuildResult build_index_gpu( cudaStream_t s, int numSms, const unsigned char* hFwd, long long n,
                                         DevBuf<U4>& oBwt, DevBuf<long long>& oSa, DevBuf<unsigned char>& oPac,
                                         int64_t& launches )
{
    const long long N = 2 * n;
    if( n <= 0 || N >= 0x7fffffffll )
        throw std::runtime_error( "index_build: forward length must be in (0, 2^30) for the 32-bit suffix sorter" );
    const int G = numSms * 8, B = 256;
    DevBuf<unsigned char> fwd, T;
    DevBuf<unsigned long long> keyA, keyB;
    DevBuf<unsigned int> idxA, idxB, rank, flag, rankSorted;
    fwd.reserve( (size_t)n );
    T.reserve( (size_t)N + MA_IB_K0 );
    keyA.reserve( (size_t)N ), keyB.reserve( (size_t)N );
    idxA.reserve( (size_t)N ), idxB.reserve( (size_t)N );
    rank.reserve( (size_t)N ), flag.reserve( (size_t)N ), rankSorted.reserve( (size_t)N );
    MA_CUDA( cudaMemcpyAsync( fwd.p, hFwd, n, cudaMemcpyHostToDevice, s ) );
    ib_text_kernel<<<G, B, 0, s>>>( fwd.p, n, T.p );
    const long long nPac = ( n + 3 ) / 4;
    oPac.reserve( (size_t)nPac + 16 );
    ib_pack_kernel<<<G, B, 0, s>>>( fwd.p, n, oPac.p, nPac );
    ib_initkeys_kernel<<<G, B, 0, s>>>( T.p, N, keyA.p, idxA.p );
    launches += 3;
    size_t tmpSort = 0, tmpScan = 0;
    cub::DeviceRadixSort::SortPairs( nullptr, tmpSort, keyA.p, keyB.p, idxA.p, idxB.p, (int)N, 0, 64, s );
    cub::DeviceScan::InclusiveSum( nullptr, tmpScan, flag.p, rankSorted.p, (int)N, s );
    DevBuf<unsigned char> tmp;
    tmp.reserve( std::max( tmpSort, tmpScan ) + 256 );
    size_t tmpBytes = tmp.cap;
    long long h = MA_IB_K0;
    int rounds = 0;
    while( true )
    {
        size_t tb = tmpBytes;
        MA_CUDA( cub::DeviceRadixSort::SortPairs( tmp.p, tb, keyA.p, keyB.p, idxA.p, idxB.p, (int)N, 0, 64, s ) );
        ib_flags_kernel<<<G, B, 0, s>>>( keyB.p, N, flag.p );
        tb = tmpBytes;
        MA_CUDA( cub::DeviceScan::InclusiveSum( tmp.p, tb, flag.p, rankSorted.p, (int)N, s ) );
        launches += 8;
        unsigned int groups = 0;
        MA_CUDA( cudaMemcpyAsync( &groups, rankSorted.p + ( N - 1 ), 4, cudaMemcpyDeviceToHost, s ) );
        MA_CUDA( cudaStreamSynchronize( s ) );
        rounds++;
        if( (long long)groups == N )
            break;
        if( h >= N )
            throw std::runtime_error( "index_build: suffixes did not become unique (internal error)" );
        ib_scatter_rank_kernel<<<G, B, 0, s>>>( idxB.p, rankSorted.p, N, rank.p );
        ib_doublekeys_kernel<<<G, B, 0, s>>>( rank.p, N, h, keyA.p, idxA.p );
        launches += 2;
        h *= 2;
    }
    const unsigned int* sa = idxB.p; // sorted suffix starts, rows 1..N
    DevBuf<long long> dPrimary;
    dPrimary.reserve( 1 );
    ib_primary_kernel<<<G, B, 0, s>>>( sa, N, dPrimary.p );
    IndexBuildResult R;
    MA_CUDA( cudaMemcpyAsync( &R.primary, dPrimary.p, 8, cudaMemcpyDeviceToHost, s ) );
    MA_CUDA( cudaStreamSynchronize( s ) );
    const long long nw = ( N + 15 ) >> 4, nblk = ( N + 127 ) / 128;
    DevBuf<unsigned int> dw;
    dw.reserve( (size_t)nw + 8 );
    ib_bwtwords_kernel<<<G, B, 0, s>>>( T.p, sa, N, R.primary, dw.p, nw );
    DevBuf<long long> cnt, excl;
    cnt.reserve( (size_t)4 * nblk ), excl.reserve( (size_t)4 * nblk );
    ib_blockcounts_kernel<<<G, B, 0, s>>>( dw.p, nw, N, nblk, cnt.p );
    size_t tmpScan2 = 0;
    cub::DeviceScan::ExclusiveSum( nullptr, tmpScan2, cnt.p, excl.p, (int)nblk, s );
    if( tmpScan2 > tmpBytes )
    {
        tmp.reserve( tmpScan2 + 256 );
        tmpBytes = tmp.cap;
    }
    for( int c = 0; c < 4; c++ )
    {
        size_t tb = tmpBytes;
        MA_CUDA( cub::DeviceScan::ExclusiveSum( tmp.p, tb, cnt.p + c * nblk, excl.p + c * nblk, (int)nblk, s ) );
    }
    R.n_words = nw + 8 * ( nblk + 1 );
    oBwt.reserve( (size_t)R.n_words / 4 + 16 );
    MA_CUDA( cudaMemsetAsync( oBwt.p, 0, ( (size_t)R.n_words / 4 + 16 ) * sizeof( U4 ), s ) );
    ib_layout_kernel<<<G, B, 0, s>>>( dw.p, nw, excl.p, cnt.p, nblk, (unsigned int*)oBwt.p );
    R.n_sa = ( N + 32 ) / 32;
    oSa.reserve( (size_t)R.n_sa + 1 );
    ib_sasample_kernel<<<G, B, 0, s>>>( sa, N, 32, oSa.p, R.n_sa );
    launches += 12;
    // totals -> L2
    long long last[ 8 ];
    for( int c = 0; c < 4; c++ )
    {
        MA_CUDA( cudaMemcpyAsync( &last[ c ], excl.p + c * nblk + nblk - 1, 8, cudaMemcpyDeviceToHost, s ) );
        MA_CUDA( cudaMemcpyAsync( &last[ 4 + c ], cnt.p + c * nblk + nblk - 1, 8, cudaMemcpyDeviceToHost, s ) );
    }
    MA_CUDA( cudaStreamSynchronize( s ) );
    MA_CUDA( cudaGetLastError( ) );
    R.L2[ 0 ] = 0;
    for( int c = 0; c < 4; c++ )
        R.L2[ c + 1 ] = R.L2[ c ] + last[ c ] + last[ 4 + c ];
    R.n_pac = nPac;
    R.rounds = rounds;
    return R;
}

// ---------------------------------------------------------------------------------------------------------------
// Large genomes (forward + reverse >= 2^31 - 1 symbols, e.g. a human-sized 3.1 Gbp reference: 6.2 G suffixes).
// The full suffix array (8 bytes x 6.2 G) and the rank arrays of prefix doubling do not fit next to each other, so the
// suffixes are sorted BUCKET BY BUCKET and only what the FM-index keeps is retained:
//   1. the text forward ++ reverse-complement as 2 bit per base, 32 bases per big-endian u64 word;
//   2. a histogram over the first MA_IB2_BIN bases assigns consecutive 10-mer bins to chunks of <= chunkCap suffixes
//      (bins are in suffix order, so chunk c holds the BWT rows that follow those of chunk c - 1);
//   3. per chunk: the positions of its suffixes are selected in one pass over the text, sorted by their first 29
//      bases (+ valid length: the shorter suffix, "$ < A", first), and the ties — suffixes that share 29 bases — are
//      refined 29 bases at a time on the tied elements only (a random genome has a handful, a repeat of length R
//      needs R / 29 rounds over the copies of that repeat);
//   4. every sorted chunk emits its BWT symbols and its SA samples (rows that are multiples of 32) and is dropped.
// The BWT of a text is unique, so the result is the one bwtLarge (fMIndex.cpp:373) computes.
#define MA_IB2_BIN 10 /* bases of the chunk-assignment histogram (4^10 bins) */

__global__ void ib2_packtext_kernel( const unsigned char* fwd, long long n, unsigned long long* P2, long long nWords )
{
    const long long N = 2 * n;
    for( long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nWords;
         w += (long long)gridDim.x * blockDim.x )
    {
        unsigned long long v = 0;
        for( int j = 0; j < 32; j++ )
        {
            const long long p = 32 * w + j;
            unsigned long long c = 0;
            if( p < N )
                c = p < n ? ( fwd[ p ] & 3u ) : ( 3u - ( fwd[ N - 1 - p ] & 3u ) );
            v |= c << ( 62 - 2 * j );
        }
        P2[ w ] = v;
    }
}

// the 32 bases starting at pos, first base in the top bits (zeros beyond the end: P2 carries two zero words of padding)
__device__ __forceinline__ unsigned long long ib2_window( const unsigned long long* __restrict__ P2, long long pos )
{
    const unsigned long long hi = P2[ pos >> 5 ], lo = P2[ ( pos >> 5 ) + 1 ];
    const int sh = (int)( pos & 31 ) * 2;
    return sh ? ( hi << sh ) | ( lo >> ( 64 - sh ) ) : hi;
}
__device__ __forceinline__ int ib2_base( const unsigned long long* __restrict__ P2, long long pos )
{
    return (int)( P2[ pos >> 5 ] >> ( 62 - 2 * (int)( pos & 31 ) ) ) & 3;
}
// sort key of the suffix starting at pos: 29 bases + 5 bits of valid length (same format as ib_initkeys_kernel)
__device__ __forceinline__ unsigned long long ib2_key( const unsigned long long* __restrict__ P2, long long N,
                                                       long long pos )
{
    if( pos >= N )
        return 0;
    const long long valid = N - pos < MA_IB_K0 ? N - pos : MA_IB_K0;
    return ( ( ib2_window( P2, pos ) >> 6 ) << 5 ) | (unsigned long long)valid;
}

__global__ void ib2_hist_kernel( const unsigned long long* P2, long long N, unsigned long long* hist )
{
    for( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N;
         i += (long long)gridDim.x * blockDim.x )
        atomicAdd( &hist[ ib2_window( P2, i ) >> ( 64 - 2 * MA_IB2_BIN ) ], 1ull );
}

// positions whose bin lies in [b0, b1) -> (key, position), in any order (one atomic per warp)
__global__ void ib2_select_kernel( const unsigned long long* P2, long long N, unsigned int b0, unsigned int b1,
                                   unsigned long long* key, unsigned long long* val, unsigned long long* cursor )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for( long long base = ( (long long)blockIdx.x * blockDim.x + threadIdx.x ) & ~31ll; base < N; base += stride )
    {
        const long long i = base + lane;
        bool in = false;
        if( i < N )
        {
            const unsigned int b = (unsigned int)( ib2_window( P2, i ) >> ( 64 - 2 * MA_IB2_BIN ) );
            in = b >= b0 && b < b1;
        }
        const unsigned m = __ballot_sync( FULL, in );
        if( m == 0 )
            continue;
        unsigned long long o = 0;
        if( lane == __ffs( m ) - 1 )
            o = atomicAdd( cursor, (unsigned long long)__popc( m ) );
        o = __shfl_sync( FULL, o, __ffs( m ) - 1 );
        if( in )
        {
            const unsigned long long slot = o + __popc( m & ( ( 1u << lane ) - 1 ) );
            key[ slot ] = ib2_key( P2, N, i );
            val[ slot ] = (unsigned long long)i;
        }
    }
}

// element j belongs to a group of equal keys (full 29 valid bases only: shorter suffixes have unique keys)
__global__ void ib2_tieflags_kernel( const unsigned long long* key, long long n, unsigned char* tied,
                                     unsigned int* head )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x )
    {
        const unsigned long long k = key[ j ];
        const bool eqPrev = j > 0 && key[ j - 1 ] == k, eqNext = j + 1 < n && key[ j + 1 ] == k;
        tied[ j ] = ( ( eqPrev || eqNext ) && ( k & 31 ) == MA_IB_K0 ) ? 1 : 0;
        head[ j ] = eqPrev ? 0u : 1u;
    }
}

// gather the tied elements: slot (index in the chunk's sorted array), group label, position
__global__ void ib2_gathertied_kernel( const unsigned int* slotIn, long long m, const unsigned int* labelAll,
                                       const unsigned long long* val, unsigned int* label, unsigned long long* pos )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (long long)gridDim.x * blockDim.x )
    {
        const unsigned int s = slotIn[ k ];
        label[ k ] = labelAll[ s ];
        pos[ k ] = val[ s ];
    }
}

__global__ void ib2_iota_kernel( unsigned int* a, long long n )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (long long)gridDim.x * blockDim.x )
        a[ k ] = (unsigned int)k;
}

__global__ void ib2_nextkeys_kernel( const unsigned long long* P2, long long N, const unsigned long long* pos,
                                     long long h, long long m, unsigned long long* key2, unsigned int* perm )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (long long)gridDim.x * blockDim.x )
    {
        key2[ k ] = ib2_key( P2, N, (long long)pos[ k ] + h );
        perm[ k ] = (unsigned int)k;
    }
}

__global__ void ib2_gatherlabel_kernel( const unsigned int* perm, const unsigned int* label, long long m,
                                        unsigned int* out )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (long long)gridDim.x * blockDim.x )
        out[ k ] = label[ perm[ k ] ];
}

// after the (label, key2) sort: write the refined order back and flag what is still tied
__global__ void ib2_refine_kernel( const unsigned int* perm, const unsigned int* labelS, const unsigned long long* key2,
                                   const unsigned long long* pos, const unsigned int* slot, long long m,
                                   unsigned long long* val, unsigned long long* posOut, unsigned char* tied,
                                   unsigned int* head )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (long long)gridDim.x * blockDim.x )
    {
        const unsigned int p = perm[ k ];
        const unsigned long long ps = pos[ p ], k2 = key2[ p ];
        val[ slot[ k ] ] = ps; // slots are increasing, labels increase with the slot: the k-th tied slot gets the k-th
        posOut[ k ] = ps;
        const bool eqPrev = k > 0 && labelS[ k - 1 ] == labelS[ k ] && key2[ perm[ k - 1 ] ] == k2;
        const bool eqNext = k + 1 < m && labelS[ k + 1 ] == labelS[ k ] && key2[ perm[ k + 1 ] ] == k2;
        tied[ k ] = ( ( eqPrev || eqNext ) && ( k2 & 31 ) == MA_IB_K0 ) ? 1 : 0;
        head[ k ] = eqPrev ? 0u : 1u;
    }
}

__global__ void ib2_compact_kernel( const unsigned int* sel, long long m2, const unsigned int* slot,
                                    const unsigned int* labelAll, const unsigned long long* pos, unsigned int* slotOut,
                                    unsigned int* labelOut, unsigned long long* posOut )
{
    for( long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < m2; k += (long long)gridDim.x * blockDim.x )
    {
        const unsigned int s = sel[ k ];
        slotOut[ k ] = slot[ s ], labelOut[ k ] = labelAll[ s ], posOut[ k ] = pos[ s ];
    }
}

// BWT symbol (one byte per row, rows 0..N; the primary row holds 0) and SA samples of the rows of one chunk
__global__ void ib2_emit_kernel( const unsigned long long* P2, const unsigned long long* val, long long cnt,
                                 long long rowBase, unsigned char* bwtRow, long long* sa, int intv, long long* primary )
{
    for( long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < cnt; j += (long long)gridDim.x * blockDim.x )
    {
        const long long row = rowBase + j, p = (long long)val[ j ];
        if( p == 0 )
            *primary = row;
        bwtRow[ row ] = p == 0 ? (unsigned char)0 : (unsigned char)ib2_base( P2, p - 1 );
        if( row % intv == 0 )
            sa[ row / intv ] = p;
    }
}

__global__ void ib2_bwtwords_kernel( const unsigned char* bwtRow, long long N, long long primary, unsigned int* dw,
                                     long long nw )
{
    for( long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x; w < nw; w += (long long)gridDim.x * blockDim.x )
    {
        unsigned int v = 0;
        for( int j = 0; j < 16; j++ )
        {
            const long long b = 16 * w + j;
            unsigned int c = 0;
            if( b < N )
                c = bwtRow[ b < primary ? b : b + 1 ];
            v |= c << ( ( 15 - j ) * 2 );
        }
        dw[ w ] = v;
    }
}

inline IndexBuildResult build_index_gpu_large( cudaStream_t s, int numSms, const unsigned char* hFwd, long long n,
                                               long long chunkCap, DevBuf<U4>& oBwt, DevBuf<long long>& oSa,
                                               DevBuf<unsigned char>& oPac, int64_t& launches )
{
    const long long N = 2 * n;
    if( n <= 0 || N >= ( 1ll << 39 ) )
        throw std::runtime_error( "index_build: forward length must be in (0, 2^38)" );
    if( chunkCap <= 0 || chunkCap > ( 1ll << 30 ) )
        chunkCap = 1ll << 29;
    const int G = numSms * 8, B = 256;
    const long long nWords = ( N + 31 ) / 32;
    DevBuf<unsigned long long> P2;
    P2.reserve( (size_t)nWords + 4 );
    const long long nPac = ( n + 3 ) / 4;
    oPac.reserve( (size_t)nPac + 16 );
    {
        DevBuf<unsigned char> fwd;
        fwd.reserve( (size_t)n );
        MA_CUDA( cudaMemcpyAsync( fwd.p, hFwd, n, cudaMemcpyHostToDevice, s ) );
        MA_CUDA( cudaMemsetAsync( P2.p + nWords, 0, 4 * sizeof( unsigned long long ), s ) );
        ib2_packtext_kernel<<<G, B, 0, s>>>( fwd.p, n, P2.p, nWords );
        ib_pack_kernel<<<G, B, 0, s>>>( fwd.p, n, oPac.p, nPac );
        MA_CUDA( cudaStreamSynchronize( s ) );
        launches += 2;
    }
    // chunk assignment
    const long long nBins = 1ll << ( 2 * MA_IB2_BIN );
    DevBuf<unsigned long long> dHist, dCursor;
    dHist.reserve( (size_t)nBins ), dCursor.reserve( 16 );
    MA_CUDA( cudaMemsetAsync( dHist.p, 0, nBins * sizeof( unsigned long long ), s ) );
    ib2_hist_kernel<<<G, B, 0, s>>>( P2.p, N, dHist.p );
    launches++;
    std::vector<unsigned long long> hist( (size_t)nBins );
    MA_CUDA( cudaMemcpyAsync( hist.data( ), dHist.p, nBins * sizeof( unsigned long long ), cudaMemcpyDeviceToHost, s ) );
    MA_CUDA( cudaStreamSynchronize( s ) );
    struct Chunk
    {
        unsigned int b0, b1;
        long long cnt;
    };
    std::vector<Chunk> chunks;
    long long maxCnt = 0;
    for( long long b = 0; b < nBins; )
    {
        long long c = 0, e = b;
        while( e < nBins && ( e == b || c + (long long)hist[ e ] <= chunkCap ) )
            c += (long long)hist[ e++ ];
        if( c > ( 1ll << 31 ) - 64 )
            throw std::runtime_error( "index_build: more than 2^31 suffixes share their first 10 bases" );
        if( c > 0 )
            chunks.push_back( Chunk{ (unsigned int)b, (unsigned int)e, c } ), maxCnt = std::max( maxCnt, c );
        b = e;
    }
    // outputs that persist over the chunks
    DevBuf<unsigned char> bwtRow;
    bwtRow.reserve( (size_t)N + 2 );
    IndexBuildResult R;
    R.n_sa = ( N + 32 ) / 32;
    oSa.reserve( (size_t)R.n_sa + 1 );
    DevBuf<long long> dPrimary;
    dPrimary.reserve( 1 );
    // chunk buffers
    DevBuf<unsigned long long> keyA, keyB, valA, valB;
    keyA.reserve( (size_t)maxCnt ), keyB.reserve( (size_t)maxCnt ), valA.reserve( (size_t)maxCnt ), valB.reserve( (size_t)maxCnt );
    DevBuf<unsigned char> tied;
    DevBuf<unsigned int> head, labelAll, slotA, slotB, labA, labB, labS, perm, permB, sel;
    DevBuf<unsigned long long> posA, posB, key2, key2B;
    tied.reserve( (size_t)maxCnt ), head.reserve( (size_t)maxCnt ), labelAll.reserve( (size_t)maxCnt );
    size_t tmpBytes = 0;
    {
        size_t a = 0, b2 = 0, c = 0;
        cub::DeviceRadixSort::SortPairs( nullptr, a, keyA.p, keyB.p, valA.p, valB.p, (int)maxCnt, 0, 64, s );
        cub::DeviceScan::InclusiveSum( nullptr, b2, head.p, labelAll.p, (int)maxCnt, s );
        cub::DeviceSelect::Flagged( nullptr, c, perm.p, tied.p, sel.p, (int*)nullptr, (int)maxCnt, s );
        tmpBytes = std::max( a, std::max( b2, c ) ) + 256;
    }
    DevBuf<unsigned char> tmp;
    tmp.reserve( tmpBytes );
    tmpBytes = tmp.cap;
    DevBuf<int> dNum;
    dNum.reserve( 4 );
    auto ensureTied = [ & ]( long long m ) {
        slotA.reserve( (size_t)m ), slotB.reserve( (size_t)m ), labA.reserve( (size_t)m ), labB.reserve( (size_t)m );
        labS.reserve( (size_t)m ), perm.reserve( (size_t)m ), permB.reserve( (size_t)m ), sel.reserve( (size_t)m );
        posA.reserve( (size_t)m ), posB.reserve( (size_t)m ), key2.reserve( (size_t)m ), key2B.reserve( (size_t)m );
    };
    MA_CUDA( cudaMemsetAsync( dPrimary.p, 0xff, 8, s ) );
    // row 0 is the suffix "$": its BWT symbol is the last base of the text, its SA sample slot holds -1
    long long rowBase = 1;
    int rounds = 0;
    for( const Chunk& ck : chunks )
    {
        const long long cnt = ck.cnt;
        MA_CUDA( cudaMemsetAsync( dCursor.p, 0, 8, s ) );
        ib2_select_kernel<<<G, B, 0, s>>>( P2.p, N, ck.b0, ck.b1, keyA.p, valA.p, dCursor.p );
        size_t tb = tmpBytes;
        MA_CUDA( cub::DeviceRadixSort::SortPairs( tmp.p, tb, keyA.p, keyB.p, valA.p, valB.p, (int)cnt, 0, 64, s ) );
        ib2_tieflags_kernel<<<G, B, 0, s>>>( keyB.p, cnt, tied.p, head.p );
        tb = tmpBytes;
        MA_CUDA( cub::DeviceScan::InclusiveSum( tmp.p, tb, head.p, labelAll.p, (int)cnt, s ) );
        // indices of the tied elements (increasing)
        ensureTied( 1024 );
        long long m = 0;
        {
            // count first so that the buffers can be sized
            DevBuf<unsigned int>& iota = permB; // scratch
            iota.reserve( (size_t)cnt );
            ib2_iota_kernel<<<G, B, 0, s>>>( iota.p, cnt );
            sel.reserve( (size_t)cnt );
            tb = tmpBytes;
            MA_CUDA( cub::DeviceSelect::Flagged( tmp.p, tb, iota.p, tied.p, sel.p, dNum.p, (int)cnt, s ) );
            int hm = 0;
            MA_CUDA( cudaMemcpyAsync( &hm, dNum.p, 4, cudaMemcpyDeviceToHost, s ) );
            MA_CUDA( cudaStreamSynchronize( s ) );
            m = hm;
        }
        launches += 8;
        if( m > 0 )
        {
            ensureTied( m );
            MA_CUDA( cudaMemcpyAsync( slotA.p, sel.p, m * 4, cudaMemcpyDeviceToDevice, s ) );
            ib2_gathertied_kernel<<<G, B, 0, s>>>( slotA.p, m, labelAll.p, valB.p, labA.p, posA.p );
            launches++;
        }
        long long h = MA_IB_K0;
        while( m > 0 )
        {
            if( h > N + MA_IB_K0 )
                throw std::runtime_error( "index_build: suffixes did not become unique (internal error)" );
            rounds++;
            // order the tied elements by (label, next 29 bases): LSD with two stable radix sorts
            ib2_nextkeys_kernel<<<G, B, 0, s>>>( P2.p, N, posA.p, h, m, key2.p, perm.p );
            tb = tmpBytes;
            MA_CUDA( cub::DeviceRadixSort::SortPairs( tmp.p, tb, key2.p, key2B.p, perm.p, permB.p, (int)m, 0, 64, s ) );
            ib2_gatherlabel_kernel<<<G, B, 0, s>>>( permB.p, labA.p, m, labB.p );
            tb = tmpBytes;
            MA_CUDA( cub::DeviceRadixSort::SortPairs( tmp.p, tb, labB.p, labS.p, permB.p, perm.p, (int)m, 0, 32, s ) );
            // perm: tied-list index in (label, key2) order; labS: its label
            ib2_refine_kernel<<<G, B, 0, s>>>( perm.p, labS.p, key2.p, posA.p, slotA.p, m, valB.p, posB.p, tied.p, head.p );
            // new labels: groups are now (old label, key2)
            tb = tmpBytes;
            MA_CUDA( cub::DeviceScan::InclusiveSum( tmp.p, tb, head.p, labelAll.p, (int)m, s ) );
            ib2_iota_kernel<<<G, B, 0, s>>>( permB.p, m );
            tb = tmpBytes;
            MA_CUDA( cub::DeviceSelect::Flagged( tmp.p, tb, permB.p, tied.p, sel.p, dNum.p, (int)m, s ) );
            int hm = 0;
            MA_CUDA( cudaMemcpyAsync( &hm, dNum.p, 4, cudaMemcpyDeviceToHost, s ) );
            MA_CUDA( cudaStreamSynchronize( s ) );
            launches += 12;
            if( hm > 0 )
            {
                ib2_compact_kernel<<<G, B, 0, s>>>( sel.p, hm, slotA.p, labelAll.p, posB.p, slotB.p, labB.p, posA.p );
                MA_CUDA( cudaMemcpyAsync( slotA.p, slotB.p, (size_t)hm * 4, cudaMemcpyDeviceToDevice, s ) );
                MA_CUDA( cudaMemcpyAsync( labA.p, labB.p, (size_t)hm * 4, cudaMemcpyDeviceToDevice, s ) );
                launches++;
            }
            m = hm;
            h += MA_IB_K0;
        }
        ib2_emit_kernel<<<G, B, 0, s>>>( P2.p, valB.p, cnt, rowBase, bwtRow.p, oSa.p, 32, dPrimary.p );
        launches++;
        rowBase += cnt;
    }
    if( rowBase != N + 1 )
        throw std::runtime_error( "index_build: chunk counts do not add up (internal error)" );
    MA_CUDA( cudaMemcpyAsync( &R.primary, dPrimary.p, 8, cudaMemcpyDeviceToHost, s ) );
    {
        // row 0: BWT symbol = last base of the text; SA sample -1
        unsigned char last = (unsigned char)( 3 - ( hFwd[ 0 ] & 3 ) );
        const long long minus1 = -1;
        MA_CUDA( cudaMemcpyAsync( bwtRow.p, &last, 1, cudaMemcpyHostToDevice, s ) );
        MA_CUDA( cudaMemcpyAsync( oSa.p, &minus1, 8, cudaMemcpyHostToDevice, s ) );
    }
    MA_CUDA( cudaStreamSynchronize( s ) );
    if( R.primary <= 0 )
        throw std::runtime_error( "index_build: primary row not found (internal error)" );
    // release the chunk buffers before the output tables are allocated
    keyA.release( ), keyB.release( ), valA.release( ), valB.release( ), P2.release( );
    const long long nw = ( N + 15 ) >> 4, nblk = ( N + 127 ) / 128;
    DevBuf<unsigned int> dw;
    dw.reserve( (size_t)nw + 8 );
    ib2_bwtwords_kernel<<<G, B, 0, s>>>( bwtRow.p, N, R.primary, dw.p, nw );
    MA_CUDA( cudaStreamSynchronize( s ) );
    bwtRow.release( );
    DevBuf<long long> cnt, excl;
    cnt.reserve( (size_t)4 * nblk ), excl.reserve( (size_t)4 * nblk );
    ib_blockcounts_kernel<<<G, B, 0, s>>>( dw.p, nw, N, nblk, cnt.p );
    size_t tmpScan2 = 0;
    cub::DeviceScan::ExclusiveSum( nullptr, tmpScan2, cnt.p, excl.p, (int)nblk, s );
    tmp.reserve( tmpScan2 + 256 );
    for( int c = 0; c < 4; c++ )
    {
        size_t tb = tmp.cap;
        MA_CUDA( cub::DeviceScan::ExclusiveSum( tmp.p, tb, cnt.p + c * nblk, excl.p + c * nblk, (int)nblk, s ) );
    }
    R.n_words = nw + 8 * ( nblk + 1 );
    oBwt.reserve( (size_t)R.n_words / 4 + 16 );
    MA_CUDA( cudaMemsetAsync( oBwt.p, 0, ( (size_t)R.n_words / 4 + 16 ) * sizeof( U4 ), s ) );
    ib_layout_kernel<<<G, B, 0, s>>>( dw.p, nw, excl.p, cnt.p, nblk, (unsigned int*)oBwt.p );
    launches += 11;
    long long last[ 8 ];
    for( int c = 0; c < 4; c++ )
    {
        MA_CUDA( cudaMemcpyAsync( &last[ c ], excl.p + c * nblk + nblk - 1, 8, cudaMemcpyDeviceToHost, s ) );
        MA_CUDA( cudaMemcpyAsync( &last[ 4 + c ], cnt.p + c * nblk + nblk - 1, 8, cudaMemcpyDeviceToHost, s ) );
    }
    MA_CUDA( cudaStreamSynchronize( s ) );
    MA_CUDA( cudaGetLastError( ) );
    R.L2[ 0 ] = 0;
    for( int c = 0; c < 4; c++ )
        R.L2[ c + 1 ] = R.L2[ c ] + last[ c ] + last[ 4 + c ];
    R.n_pac = nPac;
    R.rounds = rounds;
    return R;
}

} // namespace ma
