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

} // namespace ma
