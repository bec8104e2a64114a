// Banded exact DP, packed (sm_100a, DPX): the reference's kswcpp_inner_core with ALL its observable behaviour — the
// 16-aligned column ranges whose out-of-band cells run on stale state, every kswcpp_extz_t field, global and
// extension mode, any band — at two cells per lane and instruction.
//
// Replaces the scalar exact mode of ksw.cuh (ksw_rows<FAST = false>: 5.2 warp instructions per band cell on the
// PacBio end extensions and on the configs[4] sweep) whenever the scoring parameters cannot wrap the reference's int8
// difference arithmetic (ksw_p2_params_ok; the proof in ksw.cuh holds for ANY inputs inside the value intervals, hence
// also for the stale out-of-band cells of /root/reference/libs/kswcpp/inc/kswcpp_core.h:541-766).
//
// Mapping. One warp per problem, one anti-diagonal per pass as in the reference, lanes along the TARGET columns, a
// lane owns the column pair (t0, t0 + 1) of each 64-column chunk. The state of the recurrence lives in a per-warp
// circular window of W columns in shared memory as s16x2 words: every value is scaled by 8 and carries the candidate
// TAG of ksw_qs.cuh in its low three bits, so that the five-way maximum with its arg-max state is two
// __vimax3_s16x2, x' = max(a - z - e, -q - e) one __viaddmax_s16x2 and a continuation flag one __viaddmin_s16x2.
// The chunks of a row read only values of the previous row: they are independent instruction streams (ILP instead of
// occupancy: 28 W bytes of shared memory per warp).
//   * left neighbour (column t - 1): the other half of the own word or ONE shuffle per array;
//   * query bases: a circular window of the query in two copies (even / odd first index) so that the descending pair
//     (q[j], q[j - 1]) of a column pair is one aligned 32-bit word in every row;
//   * the H row (calcMaxScore, kswcpp_core.h:178-250) is int32 in shared memory, wrapped to int16 where the reference
//     uses int16 vectors; the lane-blocked POSITION of the row maximum is recomputed from the finished row only in the
//     rows that consume it (new maximum, or a z-drop test that can fire);
//   * traceback: the reference's byte per cell (state | continuation flags << 3), two bytes per lane and store.
#pragma once
#include "ksw_qs.cuh"

namespace ma
{

#ifdef MA_WARP_EMU
QS_DEV unsigned bx_nmask2( unsigned a, unsigned b )
{ // halves in which either code is N (>= 4 << 10)
    const unsigned la = a & 0xFFFFu, lb = b & 0xFFFFu, ha = a >> 16, hb = b >> 16;
    return ( ( la >= 0x1000u || lb >= 0x1000u ) ? 0xFFFFu : 0u ) | ( ( ha >= 0x1000u || hb >= 0x1000u ) ? 0xFFFF0000u : 0u );
}
QS_DEV int bx_max( int a, int b )
{
    return a > b ? a : b;
}
QS_DEV int bx_min( int a, int b )
{
    return a < b ? a : b;
}
#else
QS_DEV unsigned bx_nmask2( unsigned a, unsigned b )
{
    const unsigned cN = 0x10001000u;
    return __hge2_mask( __hmax2( *reinterpret_cast<__half2*>( &a ), *reinterpret_cast<__half2*>( &b ) ),
                        *reinterpret_cast<const __half2*>( &cN ) );
}
QS_DEV int bx_max( int a, int b )
{
    return max( a, b );
}
QS_DEV int bx_min( int a, int b )
{
    return min( a, b );
}
#endif

// The reference casts its six parameters to int8 (kswcpp_core.h:327-338); the kernel covers every set that survives
// that cast unchanged.
MA_HD inline bool ksw_bx_params_ok( const KswScore& P )
{
    return P.match > 0 && P.match <= 127 && P.mismatch <= 0 && P.mismatch >= -127 && P.q >= 0 && P.e >= 0 && P.q2 >= 0 &&
           P.e2 >= 0 && P.q + P.e <= 127 && P.q2 + P.e2 <= 127;
}

// Packed s16x2 constants, value * 256 + tag (built on the host: kernel parameters are constant-bank operands). Tags: the
// candidate that wins a tie carries the larger tag (left-aligned: diagonal 4, E 3, F 2, E2 1, F2 0; right-aligned: the
// later candidate wins and F2 never sets the state, kswcpp_core.h:668-699: diagonal 0, E 1, F 2, E2 3, F2 0).
struct BxK
{
    unsigned zMis, zXor, zN, cap, one, pq, pq2, tA, tB, tA2, tB2, ntA, ntB, ntA2, ntB2, nqe, nqe2, iUV, iX, iY, iX2, iY2, iS;
};

MA_HD inline unsigned ksw_bx_pk( int value, int tag )
{
    return ( ( (unsigned)value << 8 | (unsigned)tag ) & 0xFFFFu ) * 0x10001u;
}

MA_HD inline BxK ksw_bx_make_k( const KswScore& P, bool bLeft )
{
    const int tz = bLeft ? 4 : 0, ta = bLeft ? 3 : 1, tbb = 2, ta2 = bLeft ? 1 : 3, tb2 = 0;
    BxK K;
    K.zMis = ksw_bx_pk( P.mismatch, tz ), K.zXor = ksw_bx_pk( P.match, tz ) ^ K.zMis, K.zN = ksw_bx_pk( -P.e2, tz );
    K.cap = ksw_bx_pk( P.match, 7 ), K.one = 0x00010001u;
    // -(z * 256 + 1) - 1 + (q * 256 + 2) = (q - z) * 256
    K.pq = ( (unsigned)( P.q * 256 + 2 ) & 0xFFFFu ) * 0x10001u, K.pq2 = ( (unsigned)( P.q2 * 256 + 2 ) & 0xFFFFu ) * 0x10001u;
    K.tA = ksw_bx_pk( 0, ta ), K.tB = ksw_bx_pk( 0, tbb ), K.tA2 = ksw_bx_pk( 0, ta2 ), K.tB2 = ksw_bx_pk( 0, tb2 );
    K.ntA = ( (unsigned)( -ta ) & 0xFFFFu ) * 0x10001u, K.ntB = ( (unsigned)( -tbb ) & 0xFFFFu ) * 0x10001u;
    K.ntA2 = ( (unsigned)( -ta2 ) & 0xFFFFu ) * 0x10001u, K.ntB2 = ( (unsigned)( -tb2 ) & 0xFFFFu ) * 0x10001u;
    K.nqe = ksw_bx_pk( -( P.q + P.e ), 0 ), K.nqe2 = ksw_bx_pk( -( P.q2 + P.e2 ), 0 );
    K.iUV = ksw_bx_pk( -P.q - P.e, 0 ), K.iX = ksw_bx_pk( -P.q - P.e, ta ), K.iY = ksw_bx_pk( -P.q - P.e, tbb );
    K.iX2 = ksw_bx_pk( -P.q2 - P.e2, ta2 ), K.iY2 = ksw_bx_pk( -P.q2 - P.e2, tb2 ), K.iS = ksw_bx_pk( 0, tz );
    return K;
}

// index of column x >= 0 in a circular window of W columns (W need not be a power of two: the band-512 class uses 640)
template <int W> QS_DEV int bx_wc( const int x )
{
    return ( W & ( W - 1 ) ) == 0 ? ( x & ( W - 1 ) ) : (int)( (unsigned)x % (unsigned)W );
}
// word index of query pair k (k >= -W) in the query window of W / 2 words
template <int W> QS_DEV int bx_wq( const int k )
{
    return ( W & ( W - 1 ) ) == 0 ? ( k & ( W / 2 - 1 ) ) : (int)( (unsigned)( k + 64 * ( W / 2 ) ) % (unsigned)( W / 2 ) );
}

template <int W> struct KswBxSmem
{
    unsigned U[ W / 2 ], V[ W / 2 ], X[ W / 2 ], Y[ W / 2 ], X2[ W / 2 ], Y2[ W / 2 ]; // pair (t, t + 1) at [(t & (W-1)) >> 1]
    unsigned S[ W / 2 ]; // score profile of the last row that wrote the column (read stale by out-of-band cells)
    unsigned TC[ W / 2 ]; // target codes c << 10
    unsigned QE[ W / 2 ], QO[ W / 2 ]; // QE[k] = q[2k] | q[2k-1] << 16, QO[k] = q[2k+1] | q[2k] << 16 (codes c << 10)
    int H[ W ];
    int HB[ W ]; // H row of the row that holds the running maximum
};

// position of the row maximum exactly as calcMaxScore finds it (kswcpp_core.h:178-250), from the finished H row
template <int W>
QS_DEV int ksw_bx_argmax( const int* H, const int st0, const int en0, const int lane, const int SMASK )
{
    const unsigned FULL = 0xffffffffu;
    const int NONE_T = 0x7fffffff, NONE_H = (int)0x80000000;
    const int nB = en0 - st0, nV = nB & SMASK;
    int bh = NONE_H, bt = NONE_T, th = NONE_H, tt_ = NONE_T;
    for( int dt = lane; dt < nB; dt += 32 )
    {
        const int h = H[ bx_wc<W>( st0 + dt ) ];
        if( dt < nV )
        {
            if( bt == NONE_T || h > bh )
                bh = h, bt = st0 + ( dt & SMASK );
        }
        else if( tt_ == NONE_T || h > th )
            th = h, tt_ = st0 + dt;
    }
    const int Hen0 = H[ bx_wc<W>( en0 ) ];
    for( int o = 16; o >= -SMASK; o >>= 1 )
    {
        const int oh = __shfl_xor_sync( FULL, bh, o ), ot = __shfl_xor_sync( FULL, bt, o );
        if( ot != NONE_T && ( bt == NONE_T || oh > bh || ( oh == bh && ot < bt ) ) )
            bh = oh, bt = ot;
    }
    if( bt == NONE_T || !( bh > Hen0 ) )
        bh = Hen0, bt = en0;
    const int mH = __reduce_max_sync( FULL, bh );
    int max_t = __reduce_max_sync( FULL, bt );
    if( nV < nB )
    {
        const int tm = __reduce_max_sync( FULL, tt_ == NONE_T ? NONE_H : th );
        if( tm > mH )
            max_t = __reduce_min_sync( FULL, ( tt_ != NONE_T && th == tm ) ? tt_ : NONE_T );
    }
    return max_t;
}

template <typename T, T v> struct BxTag
{
    static constexpr T value = v;
};

struct BxCell
{
    unsigned un, vn, xn, yn, x2n, y2n, tbyte; // tbyte: the traceback bytes of the two cells in bytes 0 and 2
};

// max(t, 0) - (q + e) and the continuation flag (0 or 8) of one candidate pair; t = (a - (z - q)) * 256 + tag
template <bool LEFT> QS_DEV unsigned ksw_bx_gap( const unsigned t, const unsigned tg, const unsigned ntg, const unsigned nqe,
                                                 unsigned& flag )
{
    const unsigned m = __vmaxs2( t, tg );
    if( LEFT )
        flag = __viaddmin_s16x2( m, ntg, 0x00080008u ); // a - (z - q) > 0
    else // a - (z - q) >= 0: (t - tag) is a multiple of 256
        flag = __viaddmin_s16x2( __viaddmax_s16x2( t, ntg, 0xFFF8FFF8u ), 0x00080008u, 0x00080008u );
    return __vadd2( m, nqe );
}

// one pair of cells of kswcpp_core.h:640-760. xl / vl / x2l: column t - 1, uo / yo / y2o: column t (previous row).
// Every s16 half is an int8 value of the reference in its high byte: the 16-bit wrap-around of the packed adds IS the
// int8 wrap-around of _mm_add_epi8 / _mm_sub_epi8, comparisons see the wrapped values like the reference's.
template <bool LEFT>
QS_DEV BxCell ksw_bx_cell( const BxK& K, const unsigned xl, const unsigned vl, const unsigned x2l, const unsigned uo,
                           const unsigned yo, const unsigned y2o, const unsigned z0 )
{
    BxCell o;
    const unsigned a = __vadd2( xl, vl ), b = __vadd2( yo, uo ), a2 = __vadd2( x2l, vl ), b2 = __vadd2( y2o, uo );
    unsigned d, zc;
    if( LEFT )
    {
        const unsigned zt = __vimax3_s16x2( __vimax3_s16x2( z0, a, b ), a2, b2 );
        d = 0x00040004u - ( zt & 0x00070007u ); // state = 4 - tag
        zc = __vmins2( zt, K.cap );
    }
    else
    { // right-aligned: ties go to the later candidate, state 4 is never recorded (kswcpp_core.h:693-699)
        const unsigned z4 = __vmaxs2( __vimax3_s16x2( z0, a, b ), a2 );
        d = z4 & 0x00070007u;
        zc = __vmins2( __vmaxs2( z4, b2 ), K.cap );
    }
    const unsigned z1 = ( zc & 0xFF00FF00u ) | K.one, nzc = ~z1; // z * 256 + 1; -z * 256 - 2
    o.un = __vadd2( z1, ~vl ), o.vn = __vadd2( z1, ~uo ); // z - v(t - 1), z - u(t)
    const unsigned dq = __vadd2( nzc, K.pq ), dq2 = __vadd2( nzc, K.pq2 ); // (q - z) * 256
    unsigned fa, fb, fa2, fb2;
    o.xn = ksw_bx_gap<LEFT>( __vadd2( a, dq ), K.tA, K.ntA, K.nqe, fa );
    o.yn = ksw_bx_gap<LEFT>( __vadd2( b, dq ), K.tB, K.ntB, K.nqe, fb );
    o.x2n = ksw_bx_gap<LEFT>( __vadd2( a2, dq2 ), K.tA2, K.ntA2, K.nqe2, fa2 );
    o.y2n = ksw_bx_gap<LEFT>( __vadd2( b2, dq2 ), K.tB2, K.ntB2, K.nqe2, fb2 );
    o.tbyte = fb2 * 8u + ( fa2 * 4u + ( fb * 2u + ( fa + d ) ) ); // < 128 per half
    return o;
}

// One warp, one problem, all lanes return the same ez. w >= 0. tb: >= (qlen + tlen - 1) * ncol16 bytes.
template <int W, bool LEFT>
QS_DEV void ksw_bx_rows( const BxK& K, const KswScore& P, const SeqAccess& seq, const int qlen, const int tlen, const int w,
                         const int zdrop, const bool bEarlyStop, KswBxSmem<W>& sm, unsigned char* __restrict__ tb,
                         KswOut& ez )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = qs_lane( );
    const int NEG_M = -0x40000000; // removes a cell from a maximum
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, scM = P.match;
    const int T16 = ( ( tlen + 15 ) / 16 ) * 16;
    const int ncol16 = ksw_ncol16( qlen, tlen, w );
    const int iSize = qlen > tlen ? qlen : tlen;
    const bool is16 = !( (long long)iSize * P.min16 < -32768 || (long long)iSize * P.match > 32767 );
    const int SMASK = is16 ? ~7 : ~3; // SSE lanes of the reference's H vectors: 8 x int16 or 4 x int32
    const int NEG_INF = is16 ? -32768 : (int)0x80000000;
    const int nrows = qlen + tlen - 1;
    unsigned short* const sU = reinterpret_cast<unsigned short*>( sm.U );
    unsigned short* const sV = reinterpret_cast<unsigned short*>( sm.V );
    unsigned short* const sX = reinterpret_cast<unsigned short*>( sm.X );
    unsigned short* const sY = reinterpret_cast<unsigned short*>( sm.Y );
    unsigned short* const sX2 = reinterpret_cast<unsigned short*>( sm.X2 );
    unsigned short* const sY2 = reinterpret_cast<unsigned short*>( sm.Y2 );
    int inited_end = 0; // columns [0, inited_end) of the window carry reference-visible state
    int qend = -32; // query bases j < qend are staged (j outside the query: the zero padding, compares as 'A')
    bool anyN = false;
    int last_st = -1, last_en = -1;
    unsigned cells = 0; // < 2^31 rows * band
    int prevB = 0x7fffffff;
    int bR = -1, bSt0 = 0, bEn0 = 0, bT = -1; // row, band and (once resolved) position of the running maximum (H row: sm.HB)
    long long rowOff = 0; // r * ncol16
    for( int r = 0; r < nrows; ++r, rowOff += ncol16 )
    {
        // band limits (kswcpp_core.h:541-553)
        int st0 = 0, en0 = tlen - 1;
        st0 = bx_max( st0, r - qlen + 1 );
        en0 = bx_min( en0, r );
        st0 = bx_max( st0, ( r - w + 1 ) >> 1 );
        en0 = bx_min( en0, ( r + w ) >> 1 );
        if( st0 > en0 )
        {
            ez.zdropped = 1;
            break;
        }
        cells += en0 - st0 + 1;
        const int st = st0 & ~15, en = en0 | 15;
        const int sEnd = bx_min( st0 + ( ( ( en0 - st0 ) >> 4 ) + 1 ) * 16, T16 ); // score-profile end
        const int need = bx_min( bx_max( en + 1, ( sEnd + 15 ) & ~15 ), T16 );
        if( inited_end < need )
        {
            bool n = false;
            for( int idx = inited_end + 2 * lane; idx < need; idx += 64 )
            {
                const int ki = bx_wc<W>( idx ), kk = ki >> 1;
                sm.U[ kk ] = K.iUV, sm.V[ kk ] = K.iUV, sm.X[ kk ] = K.iX, sm.Y[ kk ] = K.iY;
                sm.X2[ kk ] = K.iX2, sm.Y2[ kk ] = K.iY2, sm.S[ kk ] = K.iS;
                sm.H[ ki ] = NEG_INF, sm.H[ ki + 1 ] = NEG_INF;
                const int c0 = idx < tlen ? seq.T( idx ) : 0, c1 = idx + 1 < tlen ? seq.T( idx + 1 ) : 0;
                n |= c0 >= 4 || c1 >= 4;
                sm.TC[ kk ] = ( (unsigned)c0 << 10 ) | ( (unsigned)c1 << 26 );
            }
            anyN |= __any_sync( FULL, n );
            inited_end = need;
            __syncwarp( );
        }
        while( qend <= r - st0 + 1 )
        { // the next 32 query bases: lanes 0-15 write QE, lanes 16-31 QO
            const int k = ( qend >> 1 ) + ( lane & 15 );
            const int jl = 2 * k + ( lane >> 4 );
            const int c0 = (unsigned)jl < (unsigned)qlen ? seq.Q( jl ) : 0;
            const int c1 = (unsigned)( jl - 1 ) < (unsigned)qlen ? seq.Q( jl - 1 ) : 0;
            ( lane < 16 ? sm.QE : sm.QO )[ bx_wq<W>( k ) ] = ( (unsigned)c0 << 10 ) | ( (unsigned)c1 << 26 );
            anyN |= __any_sync( FULL, c0 >= 4 || c1 >= 4 );
            qend += 32;
        }
        const int first_col = r == 0 ? -q - e : r < P.long_thres ? -e : r == P.long_thres ? P.long_diff : -e2;
        // values entering the first column from its left neighbour (:562-579), kept in the HIGH half
        unsigned cX = K.iX & 0xFFFF0000u, cX2 = K.iX2 & 0xFFFF0000u, cV = K.iUV & 0xFFFF0000u;
        if( st > 0 )
        {
            if( st - 1 >= last_st && st - 1 <= last_en )
            {
                const int kp = bx_wc<W>( st - 1 );
                cX = (unsigned)sX[ kp ] << 16, cX2 = (unsigned)sX2[ kp ] << 16, cV = (unsigned)sV[ kp ] << 16;
            }
        }
        else
            cV = (unsigned)( first_col * 256 ) << 16;
        if( en >= r && lane == 0 )
        {
            const int kr = bx_wc<W>( r );
            sY[ kr ] = (unsigned short)( K.iY & 0xFFFFu );
            sY2[ kr ] = (unsigned short)( K.iY2 & 0xFFFFu );
            sU[ kr ] = (unsigned short)( first_col * 256 );
        }
        // old H left of en0, read before the pass updates it (:194-195); row 0: H[0] = v[0] - (q + e) (:247)
        const int ken0 = bx_wc<W>( en0 );
        const int hprev = r == 0 ? -P.qe_row0 : ( en0 > 0 ? sm.H[ ken0 > 0 ? ken0 - 1 : W - 1 ] : sm.H[ ken0 ] );
        __syncwarp( );
        const unsigned nS = (unsigned)( sEnd - st0 ), nB = (unsigned)( en0 - st0 );
        const unsigned* const qw = ( r & 1 ) ? sm.QO : sm.QE; // r - t0 has the parity of r
        int m = NEG_M, hb = NEG_M;
        unsigned char* const rowp = tb + rowOff - st;
        int kbase = bx_wc<W>( st ), qb = bx_wq<W>( ( r - st ) >> 1 ); // window indices of column st / its query pair
        // NCH consecutive 64-column chunks from column `base`. The chunks of a row only read values of the previous row:
        // all loads are issued first, so that the chunks are independent instruction streams. EDGE: the chunk may hold
        // cells outside [st0, en0) (stale profile, not part of the row maximum, H left of st0 kept) or lanes beyond en.
        auto pass = [ & ]( auto edgeTag, auto nchTag, const int base ) {
            constexpr bool EDGE = decltype( edgeTag )::value;
            constexpr int NCH = decltype( nchTag )::value;
            unsigned xo[ NCH ], vo[ NCH ], x2o[ NCH ], uo[ NCH ], yo[ NCH ], y2o[ NCH ], tcp[ NCH ], qp[ NCH ], so[ NCH ];
            unsigned long long hOld[ NCH ];
            int kk[ NCH ];
#pragma unroll
            for( int c = 0; c < NCH; c++ )
            {
                const int t0 = base + 64 * c + 2 * lane;
                int kc = kbase + 64 * c + 2 * lane; // window index of column t0 (kbase: of column `base`)
                kc = kc >= W ? kc - W : kc;
                kk[ c ] = kc >> 1;
                int kq = qb - 32 * c - lane; // ... of the query pair ((r - t0) >> 1)
                kq = kq < 0 ? kq + W / 2 : kq;
                xo[ c ] = sm.X[ kk[ c ] ], vo[ c ] = sm.V[ kk[ c ] ], x2o[ c ] = sm.X2[ kk[ c ] ];
                uo[ c ] = sm.U[ kk[ c ] ], yo[ c ] = sm.Y[ kk[ c ] ], y2o[ c ] = sm.Y2[ kk[ c ] ];
                tcp[ c ] = sm.TC[ kk[ c ] ];
                qp[ c ] = qw[ kq ];
                hOld[ c ] = *reinterpret_cast<const unsigned long long*>( &sm.H[ kc ] );
                (void)t0;
                so[ c ] = EDGE ? sm.S[ kk[ c ] ] : 0u;
            }
            unsigned upx[ NCH ], upv[ NCH ], upx2[ NCH ];
#pragma unroll
            for( int c = 0; c < NCH; c++ )
            {
                upx[ c ] = __shfl_up_sync( FULL, xo[ c ], 1 ), upv[ c ] = __shfl_up_sync( FULL, vo[ c ], 1 );
                upx2[ c ] = __shfl_up_sync( FULL, x2o[ c ], 1 );
                if( lane == 0 )
                    upx[ c ] = cX, upv[ c ] = cV, upx2[ c ] = cX2;
                if( c + 1 < NCH || base + 64 * NCH <= en )
                    cX = __shfl_sync( FULL, xo[ c ], 31 ), cV = __shfl_sync( FULL, vo[ c ], 31 ),
                    cX2 = __shfl_sync( FULL, x2o[ c ], 31 );
            }
#pragma unroll
            for( int c = 0; c < NCH; c++ )
            {
                const int t0 = base + 64 * c + 2 * lane;
                // score profile (:591-616)
                unsigned z0 = ( qs_eqmask2( tcp[ c ], qp[ c ] ) & K.zXor ) ^ K.zMis;
                if( anyN )
                {
                    const unsigned nm = bx_nmask2( tcp[ c ], qp[ c ] );
                    z0 = ( K.zN & nm ) | ( z0 & ~nm );
                }
                bool in0 = true, in1 = true, act = true, keep0 = true, keep1 = true;
                if( EDGE )
                { // out-of-band cells of the aligned range keep the stale profile of the column
                    const unsigned d0 = (unsigned)( t0 - st0 ), d1 = d0 + 1u;
                    const unsigned fm = ( d0 < nS ? 0xFFFFu : 0u ) | ( d1 < nS ? 0xFFFF0000u : 0u );
                    z0 = ( z0 & fm ) | ( so[ c ] & ~fm );
                    in0 = d0 < nB, in1 = d1 < nB;
                    act = t0 <= en;
                    keep0 = t0 >= st0, keep1 = t0 + 1 >= st0;
                }
                const BxCell C = ksw_bx_cell<LEFT>( K, __byte_perm( upx[ c ], xo[ c ], 0x5432 ),
                                                    __byte_perm( upv[ c ], vo[ c ], 0x5432 ),
                                                    __byte_perm( upx2[ c ], x2o[ c ], 0x5432 ), uo[ c ], yo[ c ], y2o[ c ], z0 );
                // H row: interior columns add v to their own H (the column at en0 is set after the pass)
                int h0 = (int)( (unsigned)hOld[ c ] + (unsigned)( (int)( C.vn << 16 ) >> 24 ) );
                int h1 = (int)( (unsigned)( hOld[ c ] >> 32 ) + (unsigned)( (int)C.vn >> 24 ) );
                if( is16 )
                    h0 = (short)h0, h1 = (short)h1;
                const int hm0 = in0 ? h0 : NEG_M, hm1 = in1 ? h1 : NEG_M;
                m = bx_max( m, bx_max( hm0, hm1 ) );
                if( bEarlyStop )
                { // scM * (query rows still below the cell), see ksw.cuh
                    const int term0 = scM * ( qlen - 1 - r + t0 );
                    hb = bx_max( hb, bx_max( hm0 + term0, hm1 + term0 + scM ) );
                }
                if( act )
                {
                    const int k = kk[ c ];
                    sm.U[ k ] = C.un, sm.V[ k ] = C.vn, sm.X[ k ] = C.xn, sm.Y[ k ] = C.yn, sm.X2[ k ] = C.x2n;
                    sm.Y2[ k ] = C.y2n, sm.S[ k ] = z0;
                    // (H of a column left of the band stays: it is read once more as the left neighbour of en0 when the
                    // band has shrunk to one column; right of en0 any value does, the column is set when it enters)
                    if( !EDGE )
                        *reinterpret_cast<unsigned long long*>( &sm.H[ 2 * k ] ) =
                            (unsigned long long)(unsigned)h0 | ( (unsigned long long)(unsigned)h1 << 32 );
                    else
                    {
                        if( keep0 )
                            sm.H[ 2 * k ] = h0;
                        if( keep1 )
                            sm.H[ 2 * k + 1 ] = h1;
                    }
                    *reinterpret_cast<unsigned short*>( rowp + t0 ) = (unsigned short)__byte_perm( C.tbyte, 0, 0x4420 );
                }
            }
        };
        for( int base = st; base <= en; )
        {
            // interior chunk: every cell is in [st0, en0), hence fresh profile and part of the row maximum
            int n = 1;
            if( base < st0 || base + 63 >= en0 )
                pass( BxTag<bool, true>( ), BxTag<int, 1>( ), base );
            else if( base + 127 < en0 )
                pass( BxTag<bool, false>( ), BxTag<int, 2>( ), base ), n = 2;
            else
                pass( BxTag<bool, false>( ), BxTag<int, 1>( ), base );
            base += 64 * n, kbase += 64 * n, qb -= 32 * n;
            kbase = kbase >= W ? kbase - W : kbase, qb = qb < 0 ? qb + W / 2 : qb;
        }
        __syncwarp( );
        // the column at en0 adds u to the old H of its left neighbour (v in column 0)
        int Hen0;
        {
            const int d8 = (short)( en0 > 0 ? sU[ ken0 ] : sV[ ken0 ] );
            Hen0 = (int)( (unsigned)hprev + (unsigned)( d8 >> 8 ) );
            if( is16 )
                Hen0 = (short)Hen0;
        }
        // score-profile entries the reference writes beyond the aligned range (read, stale, by later rows)
        if( sEnd > en + 1 )
        {
            const int t0 = en + 1 + 2 * lane;
            if( t0 < sEnd )
            {
                const int kk = bx_wc<W>( t0 ) >> 1;
                const unsigned tcp = sm.TC[ kk ], qp = qw[ bx_wq<W>( ( r - t0 ) >> 1 ) ];
                unsigned z0 = ( qs_eqmask2( tcp, qp ) & K.zXor ) ^ K.zMis;
                const unsigned nm = bx_nmask2( tcp, qp );
                z0 = ( K.zN & nm ) | ( z0 & ~nm );
                const unsigned fm = 0xFFFFu | ( t0 + 1 < sEnd ? 0xFFFF0000u : 0u );
                sm.S[ kk ] = ( z0 & fm ) | ( sm.S[ kk ] & ~fm );
            }
        }
        if( lane == 0 )
            sm.H[ ken0 ] = Hen0;
        __syncwarp( );
        // the row maximum is the exact maximum of the row (only its POSITION is lane-blocked in the reference): the
        // position of a new maximum is consumed by a later z-drop test that can fire or at the end, so the H row of the
        // row that holds the running maximum is only copied aside
        const int max_H = bx_max( __reduce_max_sync( FULL, m ), Hen0 );
        if( en0 == tlen - 1 && Hen0 > ez.mte )
            ez.mte = Hen0, ez.mte_q = r - en; // sic: the aligned en
        if( r - st0 == qlen - 1 )
        {
            const int Hst0 = sm.H[ bx_wc<W>( st0 ) ];
            if( Hst0 > ez.mqe )
                ez.mqe = Hst0, ez.mqe_t = st0;
        }
        // ksw_apply_zdrop (:22-44)
        if( max_H > ez.max )
        {
            ez.max = max_H;
            bR = r, bSt0 = st0, bEn0 = en0, bT = -1;
            for( int t = ( st0 & ~1 ) + 2 * lane; t <= en0; t += 64 )
            {
                const int kt = bx_wc<W>( t );
                *reinterpret_cast<unsigned long long*>( &sm.HB[ kt ] ) = *reinterpret_cast<const unsigned long long*>( &sm.H[ kt ] );
            }
        }
        else if( zdrop >= 0 && ez.max - max_H > zdrop )
        {
            if( bR >= 0 && bT < 0 )
            { // resolved once per maximum
                __syncwarp( );
                bT = ksw_bx_argmax<W>( sm.HB, bSt0, bEn0, lane, SMASK );
            }
            const int bt = bT, bq = bR >= 0 ? bR - bT : -1;
            const int max_t = ksw_bx_argmax<W>( sm.H, st0, en0, lane, SMASK );
            if( max_t >= bt && r - max_t >= bq )
            {
                const int tl = max_t - bt, ql = ( r - max_t ) - bq;
                const int l = tl > ql ? tl - ql : ql - tl;
                if( ez.max - max_H > zdrop + l * e2 )
                {
                    ez.zdropped = 1;
                    break;
                }
            }
        }
        if( r == nrows - 1 && en0 == tlen - 1 )
            ez.score = Hen0;
        last_st = st, last_en = en;
        if( bEarlyStop )
        { // see ksw.cuh, ksw_rows
            const int B = bx_max( __reduce_max_sync( FULL, hb ), Hen0 + scM * ( qlen - 1 - r + en0 ) );
            if( r >= qlen && prevB != 0x7fffffff && bx_max( B, prevB ) <= ez.max )
            { // (the bound through query row 0 only matters once the cell bound has fallen below the maximum)
                const long long j = r + 1;
                const long long g1 = q + (long long)e * j, g2 = q2 + (long long)e2 * j;
                if( (long long)scM * qlen - ( g1 < g2 ? g1 : g2 ) <= (long long)ez.max )
                    break;
            }
            prevB = B;
        }
    }
    if( bR >= 0 )
    {
        __syncwarp( );
        ez.max_t = bT >= 0 ? bT : ksw_bx_argmax<W>( sm.HB, bSt0, bEn0, lane, SMASK ), ez.max_q = bR - ez.max_t;
    }
    ez.cells = cells;
    __syncwarp( );
}

} // namespace ma
