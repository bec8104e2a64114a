// Query-stationary register DP for the extensions that dominate the alignment path (sm_100a, DPX).
//
// Scope: the early-termination extensions of NeedlemanWunsch's read ends (needlemanWunsch.cpp:486-541: the caller
// reads only max / max_q / max_t / CIGAR) with a query of at most 192 bases, while the band of kswcpp_inner_core
// (kswcpp_core.h:541-548) is still the matrix border (rows r <= w, r < tlen). That is 99.9 % of the DP cells of the
// Illumina presets. Everything else, and every problem that leaves this regime before the early-stop bound fires
// (ksw.cuh, ksw_rows), is done by ksw_batch_kernel; a problem that has to be handed over is appended to the bin of
// that kernel (KswQsArgs::redo_*).
//
// Mapping. One warp per problem, anti-diagonal by anti-diagonal like the reference, but the lanes own QUERY
// positions instead of target columns: lane L holds the cells i = 2 (L + 32 k), 2 (L + 32 k) + 1 of block k (64
// query rows per block, NB <= 3 blocks) and keeps all DP state of those two cells in REGISTERS over the whole problem:
// the six difference values of the recurrence (kswcpp_core.h:640-760) and the H value. Per anti-diagonal a cell needs
//   * its own x, v, x2 of the previous anti-diagonal (cell (i, t-1)): registers, no memory traffic at all;
//   * u, y, y2 of query row i - 1 (cell (i-1, t)): the other half of the same register or ONE shuffle per array;
//   * its own H: H(i, t) = H(i, t-1) + u (the reference adds v to the H of column t, the same telescoping sum
//     taken along the other axis; both are the exact DP value because no int8 / int16 value can wrap in this
//     regime, ksw_p2_params_ok + ksw_qs_params_ok);
//   * the target base t = r - i: a 1 KB circular window of code pairs in shared memory.
// Nothing of the row is loaded or stored except the traceback byte.
//
// Arithmetic. Two cells per instruction as s16x2; every value is scaled by 8 and the three low bits carry a TAG
// that names the candidate (diagonal, E, F, E2, F2) with the reference's tie order, so that the arg-max state falls
// out of the max chain itself: z = __vimax3_s16x2(__vimax3_s16x2(z0, a, b), a2, b2) is the whole 5-way max WITH its
// state; x' = max(a - z - e, -q - e) is one __viaddmax_s16x2, the continuation flag one __viaddmin_s16x2 (0 or 8),
// the row maximum and the early-stop bound one __viaddmax_s16x2 each. ~27 instructions per cell pair for the
// recurrence with its traceback byte (the half2 formulation of ksw_rows_p2x2 needs 62 per pair plus 17 shared memory
// accesses).
//
// Traceback: one byte per cell at tb[r * 64 NB + i]: bits 0-2 tag (state = 4 - tag, or tag when right-aligned),
// bits 3-6 the continuation flags of kswcpp_core.h:719-757.
#pragma once
#include "ksw_types.cuh"

#ifdef MA_WARP_EMU
#define QS_DEV inline
#else
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define QS_DEV __device__ __forceinline__
#endif

namespace ma
{

#ifdef MA_WARP_EMU
QS_DEV int qs_lane( )
{
    return warpemu::lane( );
}
QS_DEV unsigned qs_prmt( unsigned a, unsigned b, unsigned s )
{
    return ma_prmt( a, b, s );
}
QS_DEV unsigned qs_eqmask2( unsigned a, unsigned b )
{
    return ma_eqmask2( a, b );
}
#else
QS_DEV int qs_lane( )
{
    return threadIdx.x & 31;
}
// prmt.b32, default mode: bit 3 of a selector nibble replicates the sign of the selected byte
QS_DEV unsigned qs_prmt( unsigned a, unsigned b, unsigned s )
{
    unsigned d;
    asm( "prmt.b32 %0, %1, %2, %3;" : "=r"( d ) : "r"( a ), "r"( b ), "r"( s ) );
    return d;
}
// base codes are kept as the half bit patterns c << 10 (0, 2^-14, ...: all distinct, none a NaN): HSET2 gives the mask
QS_DEV unsigned qs_eqmask2( unsigned a, unsigned b )
{
    return __heq2_mask( *reinterpret_cast<__half2*>( &a ), *reinterpret_cast<__half2*>( &b ) );
}
#endif

#define MA_QS_MAXQ 192 /* longest query: 3 blocks of 64 query rows */
#define MA_QS_NEG 16000 /* offset that removes a cell from a maximum without leaving int16 */

QS_DEV unsigned qs_pk( int v )
{
    return ( (unsigned)v & 0xFFFFu ) * 0x10001u;
}

// first-column value of query row i (kswcpp_core.h:562-579): v(i, -1) = H(i, -1) - H(i - 1, -1)
MA_HD inline int ksw_qs_fc( const KswScore& P, int i )
{
    return i == 0 ? -P.q - P.e : i < P.long_thres ? -P.e : i == P.long_thres ? P.long_diff : -P.e2;
}
// H(i, -1): the sum of the first-column values of rows 0..i
MA_HD inline int ksw_qs_hborder( const KswScore& P, int i )
{
    int h = -P.q - P.e;
    if( i <= 0 )
        return h;
    const int lt = P.long_thres;
    if( lt <= 0 )
        return h - P.e2 * i;
    const int n1 = ( i < lt - 1 ? i : lt - 1 ); // rows 1 .. min(i, lt - 1)
    h -= P.e * n1;
    if( i >= lt )
        h += P.long_diff - P.e2 * ( i - lt );
    return h;
}

// Does the arithmetic of this kernel hold for these scores and this band? Differences are bounded as in
// ksw_p2_params_ok (scaled by 8: < 1024); H of a cell of rows r <= w + 1 is bounded by two long gaps.
MA_HD inline bool ksw_qs_params_ok( const KswScore& P, int w )
{
    if( !ksw_p2_params_ok( P ) )
        return false;
    const int gq = P.q > P.q2 ? P.q : P.q2, ge = P.e > P.e2 ? P.e : P.e2;
    int dl = P.q + P.e - P.qe_row0;
    dl = dl < 0 ? -dl : dl;
    const long long hmin = 2ll * gq + (long long)ge * ( w + 2 + MA_QS_MAXQ ) + dl;
    const long long hmax = (long long)P.match * MA_QS_MAXQ + dl;
    return 8 * hmin + MA_QS_NEG < 32700 && 16 * hmax < 32700;
}

// 0: not for this kernel; 1..3: number of 64-row query blocks
MA_HD inline int ksw_qs_class( const KswScore& P, int qlen, int tlen, int w, int tag )
{
    if( !( tag & MA_TASK_EARLYSTOP ) || qlen < 1 || qlen > MA_QS_MAXQ || tlen < 16 || P.early_return )
        return 0;
    if( w < 0 )
        w = tlen > qlen ? tlen : qlen;
    if( w > 4096 || !( w >= qlen && ( 2 * qlen + 16 <= w || qlen + tlen - 1 <= w + 1 ) ) )
        return 0;
    if( tlen < 2 * qlen )
        return 0; // would run off the target before the bound can fire
    if( !ksw_qs_params_ok( P, w ) )
        return 0;
    return ( qlen + 63 ) / 64;
}


template <int NB> struct KswQsSmem
{
    unsigned int tp[ 256 ]; // tp[j & 255] = code(T[j]) | code(T[j - 1]) << 16, codes as c << 10
    unsigned int hbest[ 32 * NB ]; // scaled H of the row that holds the running maximum, pair (L + 32 k) at [L + 32 k]
    unsigned int hcur[ 32 * NB ]; // ... of the current row (z-drop position test)
};

struct QsK // packed s16x2 constants (built on the host: kernel parameters are constant-bank operands, not registers)
{
    unsigned zMis, zXor, cap, flA, flB, flA2, flB2, lowA, lowB, lowA2, lowB2, nlowA, nlowB, nlowA2, nlowB2, ne, ne2, one;
};

MA_HD inline unsigned ksw_qs_pk( int v )
{
    return ( (unsigned)v & 0xFFFFu ) * 0x10001u;
}

// tags: the candidate that wins a tie carries the larger tag (left-aligned: diagonal, E, F, E2, F2; right-aligned:
// the later candidate wins and F2 never sets the state, kswcpp_core.h:668-699)
MA_HD inline QsK ksw_qs_make_k( const KswScore& P, bool bLeft )
{
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, scM = P.match;
    const int tz = bLeft ? 4 : 0, ta = bLeft ? 3 : 1, tbb = 2, ta2 = bLeft ? 1 : 3, tb2 = 0;
    QsK K;
    K.zMis = ksw_qs_pk( 8 * P.mismatch + tz ), K.zXor = ksw_qs_pk( 8 * scM + tz ) ^ K.zMis, K.cap = ksw_qs_pk( 8 * scM + 7 );
    K.flA = ksw_qs_pk( 8 * ( -q - e ) + ta ), K.flB = ksw_qs_pk( 8 * ( -q - e ) + tbb );
    K.flA2 = ksw_qs_pk( 8 * ( -q2 - e2 ) + ta2 ), K.flB2 = ksw_qs_pk( 8 * ( -q2 - e2 ) + tb2 );
    const int lw = bLeft ? 0 : 8; // right-aligned: the flag is a - z >= -q, i.e. a - z - e > -q - e - 1
    K.lowA = ksw_qs_pk( 8 * ( -q - e ) + ta - lw ), K.lowB = ksw_qs_pk( 8 * ( -q - e ) + tbb - lw );
    K.lowA2 = ksw_qs_pk( 8 * ( -q2 - e2 ) + ta2 - lw ), K.lowB2 = ksw_qs_pk( 8 * ( -q2 - e2 ) + tb2 - lw );
    K.nlowA = ksw_qs_pk( -( 8 * ( -q - e ) + ta - lw ) ), K.nlowB = ksw_qs_pk( -( 8 * ( -q - e ) + tbb - lw ) );
    K.nlowA2 = ksw_qs_pk( -( 8 * ( -q2 - e2 ) + ta2 - lw ) ), K.nlowB2 = ksw_qs_pk( -( 8 * ( -q2 - e2 ) + tb2 - lw ) );
    // -e - z is formed as ~(z + 1) + (2 - e): there is no packed subtract (VIADD.16x2 only adds)
    K.ne = ksw_qs_pk( -8 * e + 2 ), K.ne2 = ksw_qs_pk( -8 * e2 + 2 ), K.one = 0x00010001u;
    return K;
}

// position of the row maximum exactly as calcMaxScore finds it (kswcpp_core.h:178-250: SSE lanes of 8 int16 / 4 int32,
// first block reaching the lane maximum, then the scalar tail, then H[en0]); Hs[i] = (scaled) H of query row i of
// anti-diagonal R, i.e. of column t = R - i
QS_DEV int ksw_qs_argmax( const short* Hs, const int R, const int st0, const int en0, const int lane, const bool is16 )
{
    const unsigned FULL = 0xffffffffu;
    const int NONE_T = 0x7fffffff, NONE_H = (int)0x80000000;
    const int SMASK = is16 ? ~7 : ~3;
    const int nB = en0 - st0, nV = nB & SMASK;
    int bh = NONE_H, bt = NONE_T, th = NONE_H, tt_ = NONE_T;
    for( int dt = lane; dt < nB; dt += 32 )
    {
        const int h = Hs[ R - ( st0 + dt ) ];
        if( dt < nV )
        {
            if( bt == NONE_T || h > bh )
                bh = h, bt = st0 + ( dt & SMASK );
        }
        else if( tt_ == NONE_T || h > th )
            th = h, tt_ = st0 + dt;
    }
    const int Hen0 = Hs[ R - en0 ];
    for( int o = 16; o >= -SMASK; o >>= 1 )
    {
        const int oh = __shfl_xor_sync( FULL, bh, o ), ot = __shfl_xor_sync( FULL, bt, o );
        if( ot != NONE_T && ( bt == NONE_T || oh > bh || ( oh == bh && ot < bt ) ) )
            bh = oh, bt = ot;
    }
    if( bt == NONE_T || !( bh > Hen0 ) )
        bh = Hen0, bt = en0;
    int mH = __reduce_max_sync( FULL, bh );
    int max_t = __reduce_max_sync( FULL, bt );
    if( nV < nB )
    {
        const int tm = __reduce_max_sync( FULL, tt_ == NONE_T ? NONE_H : th );
        if( tm > mH )
            max_t = __reduce_min_sync( FULL, ( tt_ != NONE_T && th == tm ) ? tt_ : NONE_T );
    }
    return max_t;
}

// one pair of cells. FRONT: the block holds query rows that have not entered the matrix yet (i > r): their own state
// stays at its border value. Returns the traceback bytes of the two cells in bytes 0 and 2.
template <bool LEFT, bool FRONT>
QS_DEV unsigned ksw_qs_cell( const QsK& K, unsigned& U, unsigned& V, unsigned& X, unsigned& Y, unsigned& X2,
                             unsigned& Y2, unsigned& H8, const unsigned nbU, const unsigned nbY, const unsigned nbY2,
                             const unsigned z0, const unsigned ent, unsigned& zmx )
{
    const unsigned a = __vadd2( X, V ), b = __vadd2( nbY, nbU ), a2 = __vadd2( X2, V ), b2 = __vadd2( nbY2, nbU );
    unsigned d, zc;
    if( LEFT )
    {
        const unsigned zt = __vimax3_s16x2( __vimax3_s16x2( z0, a, b ), a2, b2 );
        d = zt & 0x00070007u;
        zc = __vmins2( zt, K.cap );
        zmx = __vmaxs2( zmx, zt );
    }
    else
    { // right-aligned: ties go to the later candidate, state 4 is never recorded (kswcpp_core.h:693-699)
        const unsigned z4 = __vmaxs2( __vimax3_s16x2( z0, a, b ), a2 );
        d = z4 & 0x00070007u;
        const unsigned zt = __vmaxs2( z4, b2 );
        zc = __vmins2( zt, K.cap );
        zmx = __vmaxs2( zmx, zt );
    }
    // differences as sums of complements (a - b = a + ~b + 1 per half; there is no packed subtract):
    // z1 = z + 1 (tag bits cleared), ~z1 = -z - 2
    const unsigned z1 = ( zc & 0xFFF8FFF8u ) | K.one, nzc = ~z1;
    const unsigned un = __vadd2( z1, ~V ), vn = __vadd2( z1, ~nbU );
    const unsigned nz = __vadd2( nzc, K.ne ), nz2 = __vadd2( nzc, K.ne2 ); // -e - z
    // x' = max(a - z - e, -q - e); continuation flag a - z > -q (>= when right-aligned): 0 or 8
    unsigned xn = __viaddmax_s16x2( a, nz, K.lowA ), yn = __viaddmax_s16x2( b, nz, K.lowB );
    unsigned x2n = __viaddmax_s16x2( a2, nz2, K.lowA2 ), y2n = __viaddmax_s16x2( b2, nz2, K.lowB2 );
    const unsigned fa = __viaddmin_s16x2( xn, K.nlowA, 0x00080008u ), fb = __viaddmin_s16x2( yn, K.nlowB, 0x00080008u );
    const unsigned fa2 = __viaddmin_s16x2( x2n, K.nlowA2, 0x00080008u ),
                   fb2 = __viaddmin_s16x2( y2n, K.nlowB2, 0x00080008u );
    if( !LEFT )
        xn = __vmaxs2( xn, K.flA ), yn = __vmaxs2( yn, K.flB ), x2n = __vmaxs2( x2n, K.flA2 ),
        y2n = __vmaxs2( y2n, K.flB2 );
    const unsigned hn = __vadd2( H8, un );
    U = un, Y = yn, Y2 = y2n;
    if( FRONT )
    {
        V = ( vn & ent ) | ( V & ~ent ), X = ( xn & ent ) | ( X & ~ent ), X2 = ( x2n & ent ) | ( X2 & ~ent );
        H8 = ( hn & ent ) | ( H8 & ~ent );
    }
    else
        V = vn, X = xn, X2 = x2n, H8 = hn;
    return fb2 * 8u + ( fa2 * 4u + ( fb * 2u + ( fa + d ) ) ); // < 128 per half: no carry between the halves
}

QS_DEV int qs_hmax( unsigned v ) // larger half, sign extended
{
    const int lo = (int)(short)( v & 0xFFFFu ), hi = (int)v >> 16;
    return lo > hi ? lo : hi;
}

// One anti-diagonal r over the blocks that hold entered query rows. STEADY: r >= 64 NB - 1 (every row has entered).
// tbv[k] receives the traceback bytes of block k (bytes 0 and 2), mrow the packed row maximum.
template <int NB, bool LEFT, bool STEADY>
QS_DEV void ksw_qs_row( const QsK& K, const int r, const int lane, const int srcLane, const unsigned fcr,
                        const unsigned* __restrict__ tp, const unsigned nivec0, unsigned ( &U )[ NB ], unsigned ( &V )[ NB ],
                        unsigned ( &X )[ NB ], unsigned ( &Y )[ NB ], unsigned ( &X2 )[ NB ], unsigned ( &Y2 )[ NB ],
                        unsigned ( &H8 )[ NB ], const unsigned ( &QP )[ NB ], unsigned ( &tbv )[ NB ], unsigned& mrow,
                        unsigned& zmx )
{
    const unsigned FULL = 0xffffffffu;
    const int nbAct = STEADY ? NB : ( ( r >> 6 ) + 1 < NB ? ( r >> 6 ) + 1 : NB );
    const int kf = ( !STEADY && ( r & 63 ) != 63 && ( r >> 6 ) < NB ) ? ( r >> 6 ) : -1; // block with rows not yet entered
    const int tpi = r - 2 * lane;
    mrow = ksw_qs_pk( -2 * MA_QS_NEG );
#pragma unroll
    for( int k = NB - 1; k >= 0; --k )
    {
        if( !STEADY && k >= nbAct )
        {
            tbv[ k ] = 0;
            continue;
        }
        unsigned su = U[ k ], sy = Y[ k ], sy2 = Y2[ k ];
        if( k > 0 && lane == 31 )
            su = U[ k - 1 ], sy = Y[ k - 1 ], sy2 = Y2[ k - 1 ];
        unsigned upU = __shfl_sync( FULL, su, srcLane ), upY = __shfl_sync( FULL, sy, srcLane ),
                 upY2 = __shfl_sync( FULL, sy2, srcLane );
        if( k == 0 && lane == 0 )
            upU = fcr, upY = K.flB, upY2 = K.flB2;
        const unsigned nbU = __byte_perm( upU, U[ k ], 0x5432 ), nbY = __byte_perm( upY, Y[ k ], 0x5432 ),
                       nbY2 = __byte_perm( upY2, Y2[ k ], 0x5432 );
        const unsigned tpw = tp[ ( tpi - 64 * k ) & 255 ];
        const unsigned z0 = ( qs_eqmask2( tpw, QP[ k ] ) & K.zXor ) ^ K.zMis;
        if( !STEADY && k == kf )
        {
            const unsigned tv = __vadd2( ksw_qs_pk( r - 64 * k ), nivec0 ); // r - i per half
            const unsigned ent = ~qs_prmt( tv, 0, 0xBB99 ); // halves with i <= r
            tbv[ k ] = ksw_qs_cell<LEFT, true>( K, U[ k ], V[ k ], X[ k ], Y[ k ], X2[ k ], Y2[ k ], H8[ k ], nbU, nbY, nbY2,
                                                z0, ent, zmx );
            mrow = __vmaxs2( mrow, ( H8[ k ] & ent ) | ( ksw_qs_pk( -2 * MA_QS_NEG ) & ~ent ) );
        }
        else
        {
            tbv[ k ] = ksw_qs_cell<LEFT, false>( K, U[ k ], V[ k ], X[ k ], Y[ k ], X2[ k ], Y2[ k ], H8[ k ], nbU, nbY, nbY2,
                                                 z0, 0u, zmx );
            mrow = __vmaxs2( mrow, H8[ k ] );
        }
    }
}

// One warp, one problem; all lanes return the same result. false: the problem left the regime of this kernel (band
// term active, end of the target, an N) and must be redone by ksw_batch_kernel. Mirrors ksw_rows_p2x2 (ksw.cuh) row
// for row: same z-drop test, same two-row cadence of the early-stop bound, hence the same `cells`.
// Traceback layout: the bytes of rows r (even) and r + 1 of a cell pair share one 32-bit word:
//   tb[(r >> 1) * 128 NB + 4 (i >> 1) + 2 (r & 1) + (i & 1)].
template <int NB, bool LEFT>
QS_DEV bool ksw_qs_rows( const QsK& K, const KswScore& P, const SeqAccess& seq, const int qlen, const int tlen,
                         const int w, const int zdrop, KswQsSmem<NB>& sm, unsigned char* __restrict__ tb, KswOut& ez )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = qs_lane( );
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, scM = P.match;
    const int iSize = qlen > tlen ? qlen : tlen;
    const bool is16 = !( (long long)iSize * P.min16 < -32768 || (long long)iSize * P.match > 32767 );
    const int dl = q + e - P.qe_row0; // H[0] of row 0 uses the scalar qe from before the q / q2 swap (kswcpp_core.h:247)
    unsigned U[ NB ], V[ NB ], X[ NB ], Y[ NB ], X2[ NB ], Y2[ NB ], H8[ NB ], QP[ NB ], tbA[ NB ], tbB[ NB ];
    bool anyN = false;
#pragma unroll
    for( int k = 0; k < NB; k++ )
    {
        const int i0 = 2 * ( lane + 32 * k ), i1 = i0 + 1;
        const int c0 = i0 < qlen ? seq.Q( i0 ) : 7, c1 = i1 < qlen ? seq.Q( i1 ) : 7;
        anyN |= ( c0 >= 4 && c0 != 7 ) || ( c1 >= 4 && c1 != 7 );
        QP[ k ] = ( (unsigned)c0 << 10 ) | ( (unsigned)c1 << 26 );
        U[ k ] = 0, Y[ k ] = K.flB, Y2[ k ] = K.flB2;
        X[ k ] = K.flA, X2[ k ] = K.flA2;
        V[ k ] = ( (unsigned)( 8 * ksw_qs_fc( P, i0 ) ) & 0xFFFFu ) | ( (unsigned)( 8 * ksw_qs_fc( P, i1 ) ) << 16 );
        // query rows beyond the query (they compute the matrix of a longer query that matches nothing) start MA_QS_NEG
        // lower: they never reach a maximum
        H8[ k ] = ( (unsigned)( 8 * ( ksw_qs_hborder( P, i0 ) + dl ) - ( i0 < qlen ? 0 : MA_QS_NEG ) ) & 0xFFFFu ) |
                  ( (unsigned)( 8 * ( ksw_qs_hborder( P, i1 ) + dl ) - ( i1 < qlen ? 0 : MA_QS_NEG ) ) << 16 );
    }
    if( __any_sync( FULL, anyN ) )
        return false;
    // early-stop bound: match * (query rows still below the cell) per block, kept in registers over the whole problem
    // (the empty asm keeps the compiler from re-deriving them from the lane index in every pass)
    unsigned TERM[ NB ];
#pragma unroll
    for( int k = 0; k < NB; k++ )
    {
        TERM[ k ] = ( (unsigned)( 8 * scM * ( qlen - 1 - 2 * lane - 64 * k ) ) & 0xFFFFu ) |
                    ( (unsigned)( 8 * scM * ( qlen - 2 - 2 * lane - 64 * k ) ) << 16 );
        asm volatile( "" : "+r"( TERM[ k ] ) );
    }
    const unsigned nivec0 = ( (unsigned)( -2 * lane ) & 0xFFFFu ) | ( (unsigned)( -2 * lane - 1 ) << 16 ); // -i of block 0
    const int nrows = qlen + tlen - 1;
    const int T0 = scM * qlen;
    const unsigned fcLate = (unsigned)( -8 * e2 ) << 16; // u(-1, r) of the rows beyond the long-gap threshold
    int staged = 0;
    int lastc = 0; // code of the target base before the next staging chunk
    int ezmax8 = 0, bR = -1;
    unsigned zmx = ksw_qs_pk( -2 * MA_QS_NEG ); // the largest five-way maximum BEFORE it is clipped at the match score
    bool stop = false;
    const int srcLane = ( lane + 31 ) & 31;
    unsigned* tw = reinterpret_cast<unsigned*>( tb ) + lane; // traceback words of the current pass
    int r = 0;
    for( ; r < nrows && !stop; r += 2, tw += 32 * NB )
    {
        const bool has2 = r + 1 < nrows;
        const int rl = r + ( has2 ? 1 : 0 );
        if( rl > w || rl > tlen - 1 )
            return false; // the band term would become active / the last target column is reached
        if( staged <= r + 1 )
        { // target codes of the next 32 columns
            const int idx = staged + lane;
            const int c = idx < tlen ? seq.T( idx ) : 0;
            if( __any_sync( FULL, c >= 4 ) )
                return false;
            int pc = __shfl_up_sync( FULL, c, 1 );
            if( lane == 0 )
                pc = lastc;
            lastc = __shfl_sync( FULL, c, 31 );
            sm.tp[ idx & 255 ] = ( (unsigned)c << 10 ) | ( (unsigned)pc << 26 );
            staged += 32;
            __syncwarp( );
        }
        // u(-1, r): the value entering query row 0 from above, in the high half
        unsigned fcA = fcLate, fcB = fcLate;
        if( r <= P.long_thres || r == 0 )
            fcA = (unsigned)( 8 * ksw_qs_fc( P, r ) ) << 16, fcB = (unsigned)( 8 * ksw_qs_fc( P, r + 1 ) ) << 16;
        unsigned mrowA, mrowB = ksw_qs_pk( -2 * MA_QS_NEG );
        unsigned hb = ksw_qs_pk( -2 * MA_QS_NEG ); // bound over both rows of the pass (only its maximum is used)
        unsigned HA[ NB ]; // H of row r (position of a maximum / z-drop test of that row)
        const bool bBound = has2 && rl >= qlen;
        if( r >= 64 * NB )
        {
            ksw_qs_row<NB, LEFT, true>( K, r, lane, srcLane, fcA, sm.tp, nivec0, U, V, X, Y, X2, Y2, H8, QP, tbA, mrowA, zmx );
#pragma unroll
            for( int k = 0; k < NB; k++ )
                HA[ k ] = H8[ k ];
            if( has2 )
                ksw_qs_row<NB, LEFT, true>( K, r + 1, lane, srcLane, fcB, sm.tp, nivec0, U, V, X, Y, X2, Y2, H8, QP, tbB,
                                            mrowB, zmx );
        }
        else
        {
            ksw_qs_row<NB, LEFT, false>( K, r, lane, srcLane, fcA, sm.tp, nivec0, U, V, X, Y, X2, Y2, H8, QP, tbA, mrowA, zmx );
#pragma unroll
            for( int k = 0; k < NB; k++ )
                HA[ k ] = H8[ k ];
            if( has2 )
                ksw_qs_row<NB, LEFT, false>( K, r + 1, lane, srcLane, fcB, sm.tp, nivec0, U, V, X, Y, X2, Y2, H8, QP, tbB,
                                             mrowB, zmx );
        }
        if( !has2 )
        {
#pragma unroll
            for( int k = 0; k < NB; k++ )
                tbB[ k ] = 0;
        }
        if( bBound )
        {
#pragma unroll
            for( int k = 0; k < NB; k++ )
                hb = __viaddmax_s16x2( H8[ k ], TERM[ k ], __viaddmax_s16x2( HA[ k ], TERM[ k ], hb ) );
        }
        // traceback of the two rows: one word per cell pair
#pragma unroll
        for( int k = 0; k < NB; k++ )
            tw[ 32 * k ] = __byte_perm( tbA[ k ], tbB[ k ], 0x6420 );
        const int maxA = __reduce_max_sync( FULL, qs_hmax( mrowA ) );
        // unconditional on purpose: ptxas 12.9 predicates `has2 ? __reduce_max_sync(..) : 0` (CREDUX) with a stale
        // predicate register (profiles/r2c_ptxas_credux_predicate.md); without a second row mrowB is still -2 MA_QS_NEG
        const int maxB = __reduce_max_sync( FULL, qs_hmax( mrowB ) );
        // ksw_apply_zdrop (kswcpp_core.h:22-44) row by row, with the position resolved only when it is consumed
        for( int k2 = 0; k2 < 2; k2++ )
        {
            if( k2 == 1 && !has2 )
                break;
            const int rr = r + k2, max8 = k2 ? maxB : maxA;
            if( max8 > ezmax8 )
            {
                ezmax8 = max8, bR = rr;
                __syncwarp( );
#pragma unroll
                for( int k = 0; k < NB; k++ )
                    sm.hbest[ lane + 32 * k ] = k2 ? H8[ k ] : HA[ k ];
            }
            else if( zdrop >= 0 && ezmax8 - max8 > 8 * zdrop )
            {
                __syncwarp( );
#pragma unroll
                for( int k = 0; k < NB; k++ )
                    sm.hcur[ lane + 32 * k ] = k2 ? H8[ k ] : HA[ k ];
                __syncwarp( );
                int bt = -1, bq = -1;
                if( bR >= 0 )
                    bt = ksw_qs_argmax( reinterpret_cast<const short*>( sm.hbest ), bR, bR - qlen + 1 > 0 ? bR - qlen + 1 : 0,
                                        bR, lane, is16 ),
                    bq = bR - bt;
                const int max_t = ksw_qs_argmax( reinterpret_cast<const short*>( sm.hcur ), rr,
                                                 rr - qlen + 1 > 0 ? rr - qlen + 1 : 0, rr, lane, is16 );
                if( max_t >= bt && rr - max_t >= bq )
                {
                    const int tl = max_t - bt, ql = ( rr - max_t ) - bq;
                    const int l = tl > ql ? tl - ql : ql - tl;
                    if( ezmax8 - max8 > 8 * ( zdrop + l * e2 ) )
                    {
                        ez.zdropped = 1;
                        stop = true;
                        break;
                    }
                }
            }
        }
        if( !stop && bBound )
        { // early-stop bound over the two rows of the pass (ksw.cuh, ksw_rows_p2x2)
            const int B = __reduce_max_sync( FULL, qs_hmax( hb ) );
            if( B <= ezmax8 )
            { // (the bound through query row 0 only matters once the cell bound has fallen below the maximum)
                const int j = rl + 1;
                const int g1 = q + e * j, g2 = q2 + e2 * j;
                const int T = 8 * ( T0 - ( g1 < g2 ? g1 : g2 ) );
                if( T <= ezmax8 )
                    stop = true;
            }
        }
    }
    // band cells of the rows of all passes that were started (both rows of a pass count, as in ksw_rows_p2x2)
    const long long nR = r < nrows ? r : nrows;
    const unsigned cells =
        (unsigned)( nR <= qlen ? nR * ( nR + 1 ) / 2 : (long long)qlen * ( qlen + 1 ) / 2 + ( nR - qlen ) * qlen );
    ez.max = ezmax8 >> 3;
    if( bR >= 0 )
    {
        __syncwarp( );
        ez.max_t = ksw_qs_argmax( reinterpret_cast<const short*>( sm.hbest ), bR, bR - qlen + 1 > 0 ? bR - qlen + 1 : 0,
                                  bR, lane, is16 );
        ez.max_q = bR - ez.max_t;
    }
    ez.cells = cells;
    __syncwarp( );
    // The no-wrap argument of this kernel (ksw.cuh, packed path) needs z >= a, b, a2, b2 AND z <= match in every cell,
    // which holds as long as the clip at the match score (kswcpp_core.h:702) never changed a maximum: a consistent DP
    // never clips, an inconsistent border (negative long-gap threshold of swapped pieces) does, and then the
    // reference's int8 values can grow until they wrap. Such a problem is handed over.
    return __reduce_max_sync( FULL, qs_hmax( zmx ) ) <= 8 * scM + 7;
}

// lane 0 only: the walk of kswcpp_core.h:76-150 over this kernel's traceback layout; ops in backtrack order
QS_DEV int ksw_qs_backtrack( const unsigned char* tb, const int stride, const bool bLeft, const int qlen, const int tlen,
                             const int w, const int i0, const int j0, unsigned int* cig, const int cap )
{
    int i = i0, j = j0; // i: target, j: query
    int state = 0, n = 0;
    unsigned int cur = 0;
    auto push = [ & ]( unsigned int op, unsigned int len ) {
        if( cur != 0 && ( cur & 0xf ) == op )
            cur += len << 4;
        else
        {
            if( cur != 0 )
            {
                if( n < cap )
                    cig[ n ] = cur;
                n++;
            }
            cur = len << 4 | op;
        }
    };
    while( i >= 0 && j >= 0 )
    {
        const int r = i + j;
        const unsigned int tmp = tb[ (size_t)( r >> 1 ) * ( 2 * stride ) + 4 * ( j >> 1 ) + 2 * ( r & 1 ) + ( j & 1 ) ];
        const int tag = tmp & 7;
        const int cell = bLeft ? 4 - tag : tag;
        if( state == 0 )
            state = cell;
        else if( !( tmp >> ( state + 2 ) & 1 ) )
            state = 0;
        if( state == 0 )
            state = cell;
        if( state == 0 )
            push( 0, 1 ), --i, --j;
        else if( state == 1 || state == 3 )
            push( 2, 1 ), --i;
        else
            push( 1, 1 ), --j;
    }
    if( i >= 0 )
        push( 2, (unsigned int)i + 1 );
    if( j >= 0 )
        push( 1, (unsigned int)j + 1 );
    if( cur != 0 )
    {
        if( n < cap )
            cig[ n ] = cur;
        n++;
    }
    (void)qlen, (void)tlen, (void)w;
    return n > cap ? -1 : n;
}

} // namespace ma
