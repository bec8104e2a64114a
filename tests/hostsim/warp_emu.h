// Lock-step emulation of ONE warp on the CPU (test infrastructure): the 32 lanes of a warp-synchronous device
// function run as 32 coroutines (ucontext) that hand over at every warp-level primitive, so the same source that
// nvcc compiles for sm_100a can be checked against the oracle without a GPU (tests/hostsim/qs_sim.cpp).
// Only full-mask, convergent use of the primitives is supported; divergence at a primitive aborts.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <ucontext.h>

namespace warpemu
{
constexpr int NL = 32;
struct State
{
    ucontext_t sched;
    ucontext_t ctx[ NL ];
    char* stack[ NL ];
    bool done[ NL ];
    bool atBarrier[ NL ];
    int cur = 0;
    unsigned long long buf[ 2 ][ NL ];
    int parity[ NL ];
    std::function<void( )> fn;
};
inline State*& S( )
{
    static thread_local State* s = nullptr;
    return s;
}
inline int lane( )
{
    return S( )->cur;
}
inline void barrier( )
{
    State* s = S( );
    s->atBarrier[ s->cur ] = true;
    swapcontext( &s->ctx[ s->cur ], &s->sched );
}
inline void trampoline( )
{
    State* s = S( );
    s->fn( );
    s->done[ s->cur ] = true;
    swapcontext( &s->ctx[ s->cur ], &s->sched );
}
// runs f on 32 lanes in lock step
inline void run( std::function<void( )> f )
{
    State* s = new State( );
    S( ) = s;
    s->fn = f;
    const size_t kStack = 1 << 18;
    for( int l = 0; l < NL; l++ )
    {
        s->stack[ l ] = (char*)malloc( kStack );
        s->done[ l ] = false, s->atBarrier[ l ] = false, s->parity[ l ] = 0;
        getcontext( &s->ctx[ l ] );
        s->ctx[ l ].uc_stack.ss_sp = s->stack[ l ];
        s->ctx[ l ].uc_stack.ss_size = kStack;
        s->ctx[ l ].uc_link = nullptr;
        makecontext( &s->ctx[ l ], (void ( * )( ))trampoline, 0 );
    }
    while( true )
    {
        int nDone = 0, nBar = 0;
        for( int l = 0; l < NL; l++ )
        {
            s->cur = l;
            s->atBarrier[ l ] = false;
            swapcontext( &s->sched, &s->ctx[ l ] );
            nDone += s->done[ l ], nBar += s->atBarrier[ l ];
        }
        if( nDone == NL )
            break;
        if( nBar != NL )
        {
            fprintf( stderr, "warp_emu: divergent lanes at a warp primitive (%d at a barrier, %d finished)\n", nBar, nDone );
            abort( );
        }
    }
    for( int l = 0; l < NL; l++ )
        free( s->stack[ l ] );
    delete s;
    S( ) = nullptr;
}
template <typename T> inline T exchange( T v, int src )
{
    State* s = S( );
    const int l = s->cur, p = s->parity[ l ];
    s->parity[ l ] ^= 1;
    unsigned long long raw = 0;
    memcpy( &raw, &v, sizeof( T ) );
    s->buf[ p ][ l ] = raw;
    barrier( );
    T out;
    memcpy( &out, &s->buf[ p ][ src & 31 ], sizeof( T ) );
    return out;
}
template <typename T, typename F> inline T reduce( T v, F f )
{
    State* s = S( );
    const int l = s->cur, p = s->parity[ l ];
    s->parity[ l ] ^= 1;
    unsigned long long raw = 0;
    memcpy( &raw, &v, sizeof( T ) );
    s->buf[ p ][ l ] = raw;
    barrier( );
    T acc;
    memcpy( &acc, &s->buf[ p ][ 0 ], sizeof( T ) );
    for( int k = 1; k < NL; k++ )
    {
        T o;
        memcpy( &o, &s->buf[ p ][ k ], sizeof( T ) );
        acc = f( acc, o );
    }
    return acc;
}
} // namespace warpemu

// ---- the CUDA spellings used by the warp-synchronous kernels ------------------------------------------------
template <typename T> inline T __shfl_sync( unsigned, T v, int src )
{
    return warpemu::exchange( v, src );
}
template <typename T> inline T __shfl_up_sync( unsigned, T v, int d )
{
    const int l = warpemu::lane( );
    return warpemu::exchange( v, l - d >= 0 ? l - d : l );
}
template <typename T> inline T __shfl_xor_sync( unsigned, T v, int m )
{
    return warpemu::exchange( v, warpemu::lane( ) ^ m );
}
inline int __reduce_max_sync( unsigned, int v )
{
    return warpemu::reduce( v, []( int a, int b ) { return a > b ? a : b; } );
}
inline int __reduce_min_sync( unsigned, int v )
{
    return warpemu::reduce( v, []( int a, int b ) { return a < b ? a : b; } );
}
inline unsigned __ballot_sync( unsigned, bool pred )
{
    const unsigned bit = pred ? 1u << warpemu::lane( ) : 0u;
    return warpemu::reduce( bit, []( unsigned a, unsigned b ) { return a | b; } );
}
inline bool __any_sync( unsigned m, bool pred )
{
    return __ballot_sync( m, pred ) != 0;
}
inline void __syncwarp( )
{
    warpemu::barrier( );
}
template <typename T> inline T atomicAdd( T* p, T v )
{
    T o = *p;
    *p += v;
    return o;
}
inline int atomicExch( int* p, int v )
{
    int o = *p;
    *p = v;
    return o;
}
template <typename T> inline T atomicMax( T* p, T v )
{
    T o = *p;
    if( v > o )
        *p = v;
    return o;
}

namespace warpemu
{
inline short lo16( unsigned v )
{
    return (short)( v & 0xFFFFu );
}
inline short hi16( unsigned v )
{
    return (short)( v >> 16 );
}
inline unsigned pk16( int lo, int hi )
{
    return ( (unsigned)lo & 0xFFFFu ) | ( (unsigned)hi << 16 );
}
} // namespace warpemu
// per-half two's complement arithmetic (wraps like the hardware)
inline unsigned __vadd2( unsigned a, unsigned b )
{
    using namespace warpemu;
    return pk16( lo16( a ) + lo16( b ), hi16( a ) + hi16( b ) );
}
inline unsigned __vsub2( unsigned a, unsigned b )
{
    using namespace warpemu;
    return pk16( lo16( a ) - lo16( b ), hi16( a ) - hi16( b ) );
}
inline unsigned __vmaxs2( unsigned a, unsigned b )
{
    using namespace warpemu;
    return pk16( lo16( a ) > lo16( b ) ? lo16( a ) : lo16( b ), hi16( a ) > hi16( b ) ? hi16( a ) : hi16( b ) );
}
inline unsigned __vmins2( unsigned a, unsigned b )
{
    using namespace warpemu;
    return pk16( lo16( a ) < lo16( b ) ? lo16( a ) : lo16( b ), hi16( a ) < hi16( b ) ? hi16( a ) : hi16( b ) );
}
inline unsigned __vimax3_s16x2( unsigned a, unsigned b, unsigned c )
{
    return __vmaxs2( __vmaxs2( a, b ), c );
}
inline unsigned __vimin3_s16x2( unsigned a, unsigned b, unsigned c )
{
    return __vmins2( __vmins2( a, b ), c );
}
inline unsigned __viaddmax_s16x2( unsigned a, unsigned b, unsigned c )
{
    return __vmaxs2( __vadd2( a, b ), c );
}
inline unsigned __viaddmin_s16x2( unsigned a, unsigned b, unsigned c )
{
    return __vmins2( __vadd2( a, b ), c );
}
// prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the selected byte
inline unsigned ma_prmt( unsigned a, unsigned b, unsigned sel )
{
    const unsigned long long src = ( (unsigned long long)b << 32 ) | a;
    unsigned out = 0;
    for( int k = 0; k < 4; k++ )
    {
        const unsigned nib = ( sel >> ( 4 * k ) ) & 0xF;
        unsigned byte = (unsigned)( src >> ( 8 * ( nib & 7 ) ) ) & 0xFF;
        if( nib & 8 )
            byte = ( byte & 0x80 ) ? 0xFF : 0x00;
        out |= byte << ( 8 * k );
    }
    return out;
}
inline unsigned __byte_perm( unsigned a, unsigned b, unsigned sel )
{
    return ma_prmt( a, b, sel & 0x7777 );
}
// per-half equality mask of raw bit patterns (stands for __heq2_mask on the kernels' base codes c << 10)
inline unsigned ma_eqmask2( unsigned a, unsigned b )
{
    return ( ( a & 0xFFFFu ) == ( b & 0xFFFFu ) ? 0xFFFFu : 0u ) | ( ( a >> 16 ) == ( b >> 16 ) ? 0xFFFF0000u : 0u );
}
