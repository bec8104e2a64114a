// Banded two-piece-affine DP (extension + global) for sm_100a: one warp per DP problem, anti-diagonal wavefront.
//
// Replaces, bit-exactly, the reference's kswcpp_dispatch -> kswcpp_inner_core
//   /root/reference/libs/kswcpp/inc/kswcpp.h:165-190, kswcpp_core.h:308-841 (recurrence + traceback byte),
//   kswcpp_core.h:157-299 (H row, lane-blocked arg-max, mte/mqe), :22-44 (z-drop), :76-150 (backtrack).
//
// This file: ksw_batch_kernel (one warp per problem, dynamic task queue, window classes) and the modes it dispatches to
// in ksw_warp, most specific first:
//   ksw_rows_p2x2 (here)   in-band early-stop extensions, half2, two rows per pass; hands over if a maximum was clipped;
//   ksw_bn_rows (ksw_bn.cuh)  any problem of at most three 64-column chunks: state in registers, int8 wrap exact;
//   ksw_bx_rows (ksw_bx.cuh)  any other problem of an int8-representable score set: packed s16x2 in a shared-memory
//                             window, int8 wrap exact;
//   ksw_rows<FAST> / ksw_rows (here)  the scalar modes: score sets beyond int8, the 2048-column class.
// ksw_qs_kernel (ksw_qs.cuh) and ksw_tiny_kernel (ksw_tiny.cuh) are launched for their own bins.
//
// B200 mapping of the scalar modes:
//  * lanes run along the anti-diagonal (target index t); the reference's 16-aligned column range [st,en] is kept
//    because its out-of-band cells feed band-edge cells (they are computed from stale state on purpose);
//  * the seven int8 difference arrays + the H row live in a per-warp CIRCULAR window in shared memory
//    (W >= aligned band width + 32 columns, lazily re-initialised as the band moves right), so a 1000-column target
//    costs 1.4 KB of shared memory per warp instead of 11 KB, and 32..64 warps stay resident per SM;
//  * neighbour (t-1) state moves by __shfl_up_sync, never through memory;
//  * one traceback byte per cell is written coalesced to a per-warp slab in HBM (stays L2 resident) and walked by
//    lane 0 afterwards; the CIGAR is then copied out by the whole warp to a bump-allocated output slab.
#pragma once
#include "fmindex.cuh"
#include "ksw_types.cuh"
#include "ksw_qs.cuh"
#include "ksw_bn.cuh"
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace ma
{

__device__ __forceinline__ int w8( int x )
{
    return (int)(signed char)x;
}

// band limits of anti-diagonal r (kswcpp_core.h:541-548); returns false when out of band
__device__ __forceinline__ bool ksw_band( long long r, int qlen, int tlen, int w, int& st0, int& en0 )
{
    long long st = 0, en = tlen - 1;
    if( st < r - qlen + 1 )
        st = r - qlen + 1;
    if( en > r )
        en = r;
    if( st < ( ( r - w + 1 ) >> 1 ) )
        st = ( r - w + 1 ) >> 1;
    if( en > ( ( r + w ) >> 1 ) )
        en = ( r + w ) >> 1;
    st0 = (int)st;
    en0 = (int)en;
    return st <= en;
}

// Shared memory per warp: a circular window of W columns — 7 int8 difference arrays, the H row, the target codes of
// the window (decoded once when a column enters) — plus the query (when it fits) so that a row never touches HBM
// except for its traceback bytes.
template <int W> struct KswSmem
{
    signed char u[ W ], v[ W ], x[ W ], y[ W ], x2[ W ], y2[ W ], s[ W ];
    unsigned char tc[ W ];
    int H[ W ];
    static constexpr int QC = 2 * W <= 1024 ? 2 * W : 1024;
    unsigned char qc[ QC ];
};

// One warp, one problem. All lanes return the same KswOut (cigar_off/n_cigar are filled by the caller).
// tb: per-warp traceback slab of >= (qlen+tlen-1)*ncol16 bytes.
// Per anti-diagonal ONE fused pass over the aligned column range does: score profile, the difference recurrence with
// its traceback byte, the H-row update, the candidates of the reference's lane-blocked arg-max and the early-stop
// bound. The row loop is instantiated for left/right gap alignment and for a staged/unstaged query, runs on 32-bit
// row arithmetic and keeps all lanes on one instruction stream (out-of-range lanes compute on in-bounds shared
// memory and only their stores are predicated off).
//
// Early termination (extensions whose caller consumes only max / max_q / max_t / CIGAR): the recurrence clips z at
// the match score (kswcpp_core.h:702), so every in-band value obeys H_r[t] <= H_{r-2}[t-1] + match and no later
// cell can exceed  B = max over the last two rows of H + match * (query rows still below the cell),  nor
//   T = match * qlen - cheapest gap over r+1 target bases  for diagonals entering through query row 0.
// Once max(B_r, B_{r-1}, T_r) <= ez.max the maximum and its position are final; the reference would only go on to
// set zdropped / mqe / mte / score, which such callers never read. Validated against the reference restatement on
// adversarial problems (oracle/ksw_oracle.cpp: ma_oracle_ksw_earlystop_check, tests/test_earlystop_bound.py).
//
// FAST mode (extensions with early termination only): as long as the band term of the limits is inactive
// (rows r <= w: st0 = max(0, r-qlen+1), en0 = min(tlen-1, r)) the band edges are the matrix borders, the reference's
// out-of-band cells of the 16-aligned range feed no in-band cell and the backtrack cannot reach them, so only the
// in-band columns are computed. If such a problem is still running when the band is about to limit (r == w + 1,
// rare: the early-stop bound normally fires after ~2 * qlen rows) it is recomputed from scratch in the exact mode.
// Returns false in that case.
template <int W, bool LEFT, bool QS, bool FAST>
__device__ __forceinline__ bool ksw_rows( const KswScore& P, const SeqAccess& seq, const int qlen, const int tlen,
                                          const int w, const int zdrop, const bool bEarlyStop, KswSmem<W>& sm,
                                          unsigned char* __restrict__ tb, KswOut& ez )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int M = W - 1;
    const int NONE_T = 0x7fffffff, NONE_H = (int)0x80000000;
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, qe = q + e, qe2 = q2 + e2;
    const int T16 = ( ( tlen + 15 ) / 16 ) * 16;
    const int ncol16 = ksw_ncol16( qlen, tlen, w );
    const int iSize = qlen > tlen ? qlen : tlen;
    const bool is16 = !( (long long)iSize * P.min16 < -32768 || (long long)iSize * P.match > 32767 );
    const int SMASK = is16 ? ~7 : ~3; // SSE lanes of the reference's H vectors: 8 x int16 or 4 x int32
    const int NEG_INF = is16 ? -32768 : (int)0x80000000;
    const int init6 = w8( -q - e ), init25 = w8( -q2 - e2 );
    const int scN = -e2, scM = P.match, scX = P.mismatch;
    const int nrows = qlen + tlen - 1;
    int inited_end = 0; // columns [0, inited_end) of the circular window carry reference-visible state
    int last_st = -1, last_en = -1;
    long long cells = 0;
    int prevB = NONE_T;
    long long rowOff = 0; // r * ncol16
    for( int r = 0; r < nrows; ++r, rowOff += ncol16 )
    {
        // band limits (kswcpp_core.h:541-553)
        int st0 = 0, en0 = tlen - 1;
        st0 = max( st0, r - qlen + 1 );
        en0 = min( en0, r );
        st0 = max( st0, ( r - w + 1 ) >> 1 );
        en0 = min( en0, ( r + w ) >> 1 );
        if( st0 > en0 )
        {
            ez.zdropped = 1;
            break;
        }
        if( FAST && r > w )
            return false; // the band starts to limit: recompute in the exact mode
        cells += en0 - st0 + 1;
        const int st = st0 & ~15, en = FAST ? en0 : ( en0 | 15 );
        const int sEnd = FAST ? en0 + 1 : min( st0 + ( ( ( en0 - st0 ) >> 4 ) + 1 ) * 16, T16 ); // score-profile end
        if( FAST )
        { // only the target codes of the columns entering the band are needed
            if( inited_end <= en0 )
            {
                const int idx = inited_end + lane;
                if( idx < T16 )
                    sm.tc[ idx & M ] = idx < tlen ? (unsigned char)seq.T( idx ) : (unsigned char)0;
                inited_end += 32;
                __syncwarp( );
            }
        }
        else
        {
            const int need = min( max( en + 1, ( sEnd + 15 ) & ~15 ), T16 );
            if( inited_end < need )
            {
                for( int idx = inited_end + lane; idx < need; idx += 32 )
                {
                    const int k = idx & M;
                    sm.u[ k ] = sm.v[ k ] = sm.x[ k ] = sm.y[ k ] = (signed char)init6;
                    sm.x2[ k ] = sm.y2[ k ] = (signed char)init25;
                    sm.s[ k ] = 0;
                    sm.H[ k ] = NEG_INF;
                    sm.tc[ k ] = idx < tlen ? (unsigned char)seq.T( idx ) : (unsigned char)0;
                }
                inited_end = need;
                __syncwarp( );
            }
        }
        const int first_col =
            w8( r == 0 ? -q - e : r < P.long_thres ? -e : r == P.long_thres ? P.long_diff : -e2 );
        int cx = init6, cx2 = init25, cv = init6; // values entering the first column from its left neighbour (:562-579)
        const int c0 = FAST ? st0 : st; // first column of the pass
        if( c0 > 0 )
        {
            if( FAST || ( st - 1 >= last_st && st - 1 <= last_en ) )
                cx = sm.x[ ( c0 - 1 ) & M ], cx2 = sm.x2[ ( c0 - 1 ) & M ], cv = sm.v[ ( c0 - 1 ) & M ];
        }
        else
            cv = first_col;
        if( ( FAST ? en0 == r : en >= r ) && lane == 0 )
        {
            sm.y[ r & M ] = (signed char)init6;
            sm.y2[ r & M ] = (signed char)init25;
            sm.u[ r & M ] = (signed char)first_col;
        }
        // old H left of en0, read before the pass updates it (:194-195); row 0: H[0] = v[0] - (q + e) (:247)
        const int hprev = r == 0 ? -P.qe_row0 : ( en0 > 0 ? sm.H[ ( en0 - 1 ) & M ] : sm.H[ en0 & M ] );
        __syncwarp( );
        const int en1 = st0 + ( ( en0 - st0 ) & SMASK );
        int bh = NONE_H, bt = NONE_T; // this lane's SSE-lane candidate: first block reaching the lane maximum
        int th = NONE_H, tt_ = NONE_T; // candidate among the scalar tail [en1, en0)
        int hb = NONE_H; // early-stop bound of this lane
        const int hbBase = scM * ( qlen - 1 - r );
        int Hen0l = 0, Hst0l = 0;
        unsigned char* rowp = tb + rowOff - st;
        const unsigned nS = (unsigned)( sEnd - st0 ), nB = (unsigned)( en0 - st0 ), nV = (unsigned)( en1 - st0 );
        for( int base = c0; base <= en; base += 32 )
        {
            const int t = base + lane;
            const bool act = t <= en;
            const int k = t & M;
            const int xo = sm.x[ k ], vo = sm.v[ k ], x2o = sm.x2[ k ], ut = sm.u[ k ], yo = sm.y[ k ], y2o = sm.y2[ k ];
            const int hOld = sm.H[ k ];
            int z = FAST ? 0 : (int)sm.s[ k ]; // out-of-band cell of the aligned range: stale profile (reference)
            const unsigned dt = (unsigned)( t - st0 );
            if( FAST || dt < nS )
            { // score profile (:591-616); N scores -e2; beyond the sequences the zero padding compares as 'A'
                const int a = sm.tc[ k ];
                const int qi = r - t;
                int b = 0;
                if( (unsigned)qi < (unsigned)qlen )
                    b = QS ? (int)sm.qc[ qi ] : seq.Q( qi );
                z = ( a == 4 || b == 4 ) ? scN : ( a == b ? scM : scX );
                if( !FAST && act )
                    sm.s[ k ] = (signed char)z;
            }
            int xt1 = __shfl_up_sync( FULL, xo, 1 ), vt1 = __shfl_up_sync( FULL, vo, 1 ),
                x2t1 = __shfl_up_sync( FULL, x2o, 1 );
            if( lane == 0 )
                xt1 = cx, vt1 = cv, x2t1 = cx2;
            if( base + 32 <= en )
                cx = __shfl_sync( FULL, xo, 31 ), cv = __shfl_sync( FULL, vo, 31 ), cx2 = __shfl_sync( FULL, x2o, 31 );
            int a = w8( xt1 + vt1 ), b = w8( yo + ut ), a2 = w8( x2t1 + vt1 ), b2 = w8( y2o + ut );
            int d;
            if( LEFT )
            {
                d = a > z ? 1 : 0;
                z = max( z, a );
                d = b > z ? 2 : d;
                z = max( z, b );
                d = a2 > z ? 3 : d;
                z = max( z, a2 );
                d = b2 > z ? 4 : d;
                z = max( z, b2 );
            }
            else
            { // right-aligned: ties go to the gap, state 4 is never recorded (:693-699)
                d = z > a ? 0 : 1;
                z = max( z, a );
                d = z > b ? d : 2;
                z = max( z, b );
                d = z > a2 ? d : 3;
                z = max( z, a2 );
                z = max( z, b2 );
            }
            z = min( z, scM );
            const int un = w8( z - vt1 ), vn = w8( z - ut );
            int tmp = w8( z - q );
            a = w8( a - tmp ), b = w8( b - tmp );
            tmp = w8( z - q2 );
            a2 = w8( a2 - tmp ), b2 = w8( b2 - tmp );
            if( LEFT )
            {
                d |= a > 0 ? 0x08 : 0;
                d |= b > 0 ? 0x10 : 0;
                d |= a2 > 0 ? 0x20 : 0;
                d |= b2 > 0 ? 0x40 : 0;
            }
            else
            {
                d |= a >= 0 ? 0x08 : 0;
                d |= b >= 0 ? 0x10 : 0;
                d |= a2 >= 0 ? 0x20 : 0;
                d |= b2 >= 0 ? 0x40 : 0;
            }
            // H row (calcMaxScore, :178-250): interior columns add v, the last column adds u to its left neighbour
            const bool inB = dt < nB, isEn = t == en0;
            int h = (int)( (unsigned)( isEn ? hprev : hOld ) + (unsigned)( ( isEn && en0 > 0 ) ? un : vn ) );
            if( is16 )
                h = (short)h;
            if( act )
            {
                sm.u[ k ] = (signed char)un;
                sm.v[ k ] = (signed char)vn;
                sm.x[ k ] = (signed char)( max( a, 0 ) - qe );
                sm.y[ k ] = (signed char)( max( b, 0 ) - qe );
                sm.x2[ k ] = (signed char)( max( a2, 0 ) - qe2 );
                sm.y2[ k ] = (signed char)( max( b2, 0 ) - qe2 );
                rowp[ t ] = (unsigned char)d;
            }
            if( inB || isEn )
            {
                sm.H[ k ] = h;
                hb = max( hb, h + hbBase + scM * t );
                if( isEn )
                    Hen0l = h;
                if( t == st0 )
                    Hst0l = h;
                if( inB )
                {
                    if( dt < nV )
                    { // strict '>' keeps the first block of this SSE lane that reaches its maximum
                        if( bt == NONE_T || h > bh )
                            bh = h, bt = st0 + ( (int)dt & SMASK );
                    }
                    else if( tt_ == NONE_T || h > th )
                        th = h, tt_ = t;
                }
            }
        }
        // score-profile entries the reference writes beyond the aligned range (read, stale, by later rows)
        if( !FAST && sEnd > en + 1 )
            for( int t = en + 1 + lane; t < sEnd; t += 32 )
            {
                const int k = t & M;
                const int a = sm.tc[ k ];
                const int qi = r - t;
                int b = 0;
                if( (unsigned)qi < (unsigned)qlen )
                    b = QS ? (int)sm.qc[ qi ] : seq.Q( qi );
                sm.s[ k ] = (signed char)( ( a == 4 || b == 4 ) ? scN : ( a == b ? scM : scX ) );
            }
        const int Hen0 = __shfl_sync( FULL, Hen0l, ( en0 - c0 ) & 31 );
        // the row maximum is the exact maximum of the row (only its POSITION is lane-blocked in the reference)
        int max_H = __reduce_max_sync( FULL, max( max( bt == NONE_T ? NONE_H : bh, tt_ == NONE_T ? NONE_H : th ), Hen0 ) );
        int max_t = en0;
        // the position is consumed only by a new maximum or by a z-drop test that can fire (l >= 0)
        if( max_H > ez.max || ( zdrop >= 0 && ez.max - max_H > zdrop ) )
        {
            // lanes with equal ((t - st0) % SIZE) form one SSE lane of the reference
            const int rot = FAST ? 0 : ( ( st0 - st ) & 31 ); // in the exact mode lane % SIZE == (t - st) % SIZE
            (void)rot;
            for( int o = 16; o >= -SMASK; o >>= 1 )
            {
                const int oh = __shfl_xor_sync( FULL, bh, o ), ot = __shfl_xor_sync( FULL, bt, o );
                if( ot != NONE_T && ( bt == NONE_T || oh > bh || ( oh == bh && ot < bt ) ) )
                    bh = oh, bt = ot;
            }
            if( bt == NONE_T || !( bh > Hen0 ) )
                bh = Hen0, bt = en0;
            int mH = __reduce_max_sync( FULL, bh );
            max_t = __reduce_max_sync( FULL, bt );
            if( en1 < en0 )
            { // scalar tail [en1, en0): the first index of the tail maximum, if it beats the vector result
                const int tm = __reduce_max_sync( FULL, tt_ == NONE_T ? NONE_H : th );
                if( tm > mH )
                {
                    mH = tm;
                    max_t = __reduce_min_sync( FULL, ( tt_ != NONE_T && th == tm ) ? tt_ : NONE_T );
                }
            }
        }
        __syncwarp( );
        if( en0 == tlen - 1 && Hen0 > ez.mte )
            ez.mte = Hen0, ez.mte_q = r - en; // sic: the aligned en
        if( r - st0 == qlen - 1 )
        {
            const int Hst0 = __shfl_sync( FULL, Hst0l, ( st0 - c0 ) & 31 );
            if( Hst0 > ez.mqe )
                ez.mqe = Hst0, ez.mqe_t = st0;
        }
        // ksw_apply_zdrop (:22-44)
        if( max_H > ez.max )
            ez.max = max_H, ez.max_t = max_t, ez.max_q = r - max_t;
        else if( max_t >= ez.max_t && r - max_t >= ez.max_q )
        {
            const int tl = max_t - ez.max_t, ql = ( r - max_t ) - ez.max_q;
            const int l = tl > ql ? tl - ql : ql - tl;
            if( zdrop >= 0 && ez.max - max_H > zdrop + l * e2 )
            {
                ez.zdropped = 1;
                break;
            }
        }
        if( r == nrows - 1 && en0 == tlen - 1 )
            ez.score = Hen0;
        last_st = st, last_en = en;
        if( bEarlyStop )
        {
            const int B = __reduce_max_sync( FULL, hb );
            if( r >= qlen && prevB != NONE_T )
            {
                const long long j = r + 1;
                const long long g1 = q + (long long)e * j, g2 = q2 + (long long)e2 * j;
                const long long T = (long long)scM * qlen - ( g1 < g2 ? g1 : g2 );
                const long long bound = max( (long long)max( B, prevB ), T );
                if( bound <= (long long)ez.max )
                    break;
            }
            prevB = B;
        }
    }
    ez.cells = cells;
    __syncwarp( );
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Packed fast path: TWO cells per lane and instruction (half2 arithmetic), 64 columns per warp pass.
//
// Applies to the problems that dominate the pipeline: extensions with early termination (FAST mode above: only
// in-band cells matter) in the reference's int16 score mode, with scoring parameters for which the int8 difference
// arithmetic of the reference can never wrap. For every cell of the recurrence (kswcpp_core.h:640-760)
//     x' in [-q-e, -e],  u' = z - v >= x >= -Q,  u' <= match + Q   (Q = max(q+e, q2+e2); same for v, y, x2, y2)
// holds for any inputs inside these intervals as long as the cell's maximum z (>= a = x + v by construction) also is
// <= match WITHOUT the clip of kswcpp_core.h:702, so by induction all stored values and all intermediates stay within
// +-(2Q + match + max(|mismatch|, e2) + max(q, q2)). A consistent DP never clips (adding a base to each sequence
// raises the best score by at most one match); an inconsistent first column does (negative long-gap threshold of
// swapped pieces with e < e2, tests/golden/ksw_golden_swapped.npz): there x' = a - z - e grows from cell to cell until
// the reference's int8 wraps. The packed kernels therefore track the largest UNCLIPPED maximum over their in-band cells
// and hand a problem that ever clipped over to the banded exact mode (ksw_bx.cuh), which wraps like the reference.
// If that bound is <= 127 (ksw_p2_params_ok) the int8 wrap-around is the identity and small-integer half
// arithmetic (exact up to 2048) gives the same numbers: HADD2 / HMNMX2 / HSET2 masks work on two cells at once, the
// six difference arrays and the target / reversed-query codes are kept as halves in shared memory so that one
// 32-bit LDS/STS moves a pair, and the int16 H row is updated with VIADD.16x2 (wraps exactly like the reference).
// Pairs are aligned on even columns; the dead cell left of an odd st0 and the cell right of en0 are computed and
// stored too: both columns are rewritten before any in-band cell reads them (u/y/y2 of column r+1 are initialised
// at the start of row r+1, x/v/x2/H of a column are only read after the column's first own pass).
// The lane-blocked position of the row maximum (calcMaxScore, :178-250) is NOT tracked per cell: it is recomputed
// from the finished H row only in the rows that consume it (new maximum, or a z-drop test that can fire).
// mqe / mte / score are not produced (never read by early-stop callers, see ksw_rows).

__device__ __forceinline__ unsigned h2u( __half2 h )
{
    return *reinterpret_cast<unsigned*>( &h );
}
__device__ __forceinline__ __half2 u2h( unsigned u )
{
    return *reinterpret_cast<__half2*>( &u );
}
__device__ __forceinline__ __half2 h2i( int v )
{
    return __half2half2( __int2half_rn( v ) );
}
// (m ? a : b) per 16-bit half, m = 0xFFFF / 0 per half
__device__ __forceinline__ unsigned sel2( unsigned m, unsigned a, unsigned b )
{
    return ( a & m ) | ( b & ~m );
}

// position of the row maximum exactly as calcMaxScore finds it (SSE lanes of 8 int16, first block reaching the
// lane maximum, then the scalar tail, then H[en0]), from the finished H row in shared memory
__device__ __forceinline__ int ksw_p2_argmax( const short* __restrict__ H, const int M, const int st0, const int en0,
                                              const int lane )
{
    const unsigned FULL = 0xffffffffu;
    const int NONE_T = 0x7fffffff, NONE_H = (int)0x80000000;
    const int nB = en0 - st0, nV = nB & ~7;
    int bh = NONE_H, bt = NONE_T, th = NONE_H, tt_ = NONE_T;
    for( int dt = lane; dt < nB; dt += 32 )
    {
        const int h = H[ ( st0 + dt ) & M ];
        if( dt < nV )
        {
            if( bt == NONE_T || h > bh )
                bh = h, bt = st0 + ( dt & ~7 );
        }
        else if( tt_ == NONE_T || h > th )
            th = h, tt_ = st0 + dt;
    }
    const int Hen0 = H[ en0 & M ];
    for( int o = 16; o >= 8; o >>= 1 )
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

// ---------------------------------------------------------------------------------------------------------------
// Packed path, TWO anti-diagonals per pass. A lane keeps its column pair in registers over rows r and r + 1: the
// state arrays are loaded and stored once per two rows, the left neighbour of row r + 1 comes from a second shuffle
// of the values just computed, and the per-row bookkeeping (band, staging, reductions, bound test) runs once per
// pass. Row r + 1 enters column r + 1 (if the band still grows): its u / y / y2 start values are injected in
// registers. Three rotating H buffers: the input row, the two output rows; the buffer of the row that holds the
// running maximum is never overwritten (deferred arg-max, see above).
template <int W> struct KswSmemQ
{
    __half u[ W ], v[ W ], x[ W ], y[ W ], x2[ W ], y2[ W ], tc[ W ];
    short H[ 3 ][ W ];
    __half qa[ W ], qb[ W ]; // qa[j + 2] = code of q[qlen-1-j]; qb[j] = qa[j + 1]
};

struct P2Cell
{
    __half2 un, vn, xn, yn, x2n, y2n;
    __half2 zt; // the five-way maximum before it is clipped at the match score
    unsigned d;
};

struct P2Const
{
    __half2 hMatch, hNegQ, hNegQ2, hNegQE, hNegQE2, hE, hE2;
};

// one pair of cells of the recurrence (kswcpp_core.h:640-760) on small-integer halves
template <bool LEFT>
__device__ __forceinline__ P2Cell p2_cell( const P2Const& K, const __half2 xt1, const __half2 vt1, const __half2 x2t1,
                                           const __half2 ut, const __half2 yo, const __half2 y2o, __half2 z )
{
    P2Cell o;
    const __half2 a = __hadd2( xt1, vt1 ), b = __hadd2( yo, ut ), a2 = __hadd2( x2t1, vt1 ), b2 = __hadd2( y2o, ut );
    unsigned d;
    if( LEFT )
    {
        d = __hgt2_mask( a, z ) & 0x00010001u;
        z = __hmax2( z, a );
        d = sel2( __hgt2_mask( b, z ), 0x00020002u, d );
        z = __hmax2( z, b );
        d = sel2( __hgt2_mask( a2, z ), 0x00030003u, d );
        z = __hmax2( z, a2 );
        d = sel2( __hgt2_mask( b2, z ), 0x00040004u, d );
        z = __hmax2( z, b2 );
    }
    else
    { // right-aligned: ties go to the gap, state 4 is never recorded (:693-699)
        d = __hge2_mask( a, z ) & 0x00010001u;
        z = __hmax2( z, a );
        d = sel2( __hge2_mask( b, z ), 0x00020002u, d );
        z = __hmax2( z, b );
        d = sel2( __hge2_mask( a2, z ), 0x00030003u, d );
        z = __hmax2( z, a2 );
        z = __hmax2( z, b2 );
    }
    o.zt = z;
    z = __hmin2( z, K.hMatch );
    o.un = __hsub2( z, vt1 ), o.vn = __hsub2( z, ut );
    // x' = max(a - (z - q), 0) - (q + e) = max(a - z - e, -q - e); the continuation flag is a - z > -q
    const __half2 az = __hsub2( a, z ), bz = __hsub2( b, z ), a2z = __hsub2( a2, z ), b2z = __hsub2( b2, z );
    if( LEFT )
    {
        d |= __hgt2_mask( az, K.hNegQ ) & 0x00080008u;
        d |= __hgt2_mask( bz, K.hNegQ ) & 0x00100010u;
        d |= __hgt2_mask( a2z, K.hNegQ2 ) & 0x00200020u;
        d |= __hgt2_mask( b2z, K.hNegQ2 ) & 0x00400040u;
    }
    else
    {
        d |= __hge2_mask( az, K.hNegQ ) & 0x00080008u;
        d |= __hge2_mask( bz, K.hNegQ ) & 0x00100010u;
        d |= __hge2_mask( a2z, K.hNegQ2 ) & 0x00200020u;
        d |= __hge2_mask( b2z, K.hNegQ2 ) & 0x00400040u;
    }
    o.xn = __hmax2( __hsub2( az, K.hE ), K.hNegQE );
    o.yn = __hmax2( __hsub2( bz, K.hE ), K.hNegQE );
    o.x2n = __hmax2( __hsub2( a2z, K.hE2 ), K.hNegQE2 );
    o.y2n = __hmax2( __hsub2( b2z, K.hE2 ), K.hNegQE2 );
    o.d = d;
    return o;
}

// H pair of a row: interior cells add v to their own old H, the cell at en0 adds u (v in row 0) to hleft
__device__ __forceinline__ unsigned p2_hrow( const unsigned me, const unsigned meEn, const __half2 un, const __half2 vn,
                                             const unsigned hOwn, const unsigned hLeft )
{
    const unsigned add = h2u( __hadd2( u2h( sel2( me & meEn, h2u( un ), h2u( vn ) ) ), u2h( 0x66006600u ) ) ) &
                         0x03FF03FFu; // 1536 + value: the mantissa holds 512 + value
    return __vsub2( __vadd2( sel2( me, hLeft, hOwn ), add ), 0x02000200u );
}

template <int W, bool LEFT>
__device__ __forceinline__ bool ksw_rows_p2x2( const KswScore& P, const SeqAccess& seq, const int qlen, const int tlen,
                                               const int w, const int zdrop, KswSmemQ<W>& sm,
                                               unsigned char* __restrict__ tb, KswOut& ez )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int M = W - 1;
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, qe = q + e;
    const int scM = P.match;
    const int ncol16 = ksw_ncol16( qlen, tlen, w );
    const int nrows = qlen + tlen - 1;
    P2Const K;
    K.hMatch = h2i( scM ), K.hNegQ = h2i( -q ), K.hNegQ2 = h2i( -q2 ), K.hNegQE = h2i( -q - e );
    K.hNegQE2 = h2i( -q2 - e2 ), K.hE = h2i( e ), K.hE2 = h2i( e2 );
    const __half2 hCodeN = u2h( 0x10001000u ); // base code c is stored as the half with bits c << 10
    const unsigned uMatch = h2u( K.hMatch ), uMis = h2u( h2i( P.mismatch ) ), uN = h2u( h2i( -e2 ) );
    const unsigned short init6 = __half_as_ushort( __int2half_rn( -q - e ) ),
                         init25 = __half_as_ushort( __int2half_rn( -q2 - e2 ) );
    const unsigned init6x2 = (unsigned)init6 * 0x10001u, init25x2 = (unsigned)init25 * 0x10001u;
    // first-column values of the rows (kswcpp_core.h:562-579): row 0, rows below / at / above the long-gap threshold
    const unsigned short fc0 = init6, fc1 = __half_as_ushort( __int2half_rn( -e ) ),
                         fc2 = __half_as_ushort( __int2half_rn( P.long_diff ) ),
                         fc3 = __half_as_ushort( __int2half_rn( -e2 ) );
    unsigned* const pu = reinterpret_cast<unsigned*>( sm.u );
    unsigned* const pv = reinterpret_cast<unsigned*>( sm.v );
    unsigned* const px = reinterpret_cast<unsigned*>( sm.x );
    unsigned* const py = reinterpret_cast<unsigned*>( sm.y );
    unsigned* const px2 = reinterpret_cast<unsigned*>( sm.x2 );
    unsigned* const py2 = reinterpret_cast<unsigned*>( sm.y2 );
    unsigned* const ptc = reinterpret_cast<unsigned*>( sm.tc );
    unsigned short* const su = reinterpret_cast<unsigned short*>( sm.u );
    unsigned short* const sv = reinterpret_cast<unsigned short*>( sm.v );
    unsigned short* const sx = reinterpret_cast<unsigned short*>( sm.x );
    unsigned short* const sy = reinterpret_cast<unsigned short*>( sm.y );
    unsigned short* const sx2 = reinterpret_cast<unsigned short*>( sm.x2 );
    unsigned short* const sy2 = reinterpret_cast<unsigned short*>( sm.y2 );
    unsigned short* const stc = reinterpret_cast<unsigned short*>( sm.tc );
    {
        unsigned short* const qa = reinterpret_cast<unsigned short*>( sm.qa );
        unsigned short* const qb = reinterpret_cast<unsigned short*>( sm.qb );
        for( int j = lane; j < W; j += 32 )
        { // qa[j] = rev[j-2], qb[j] = qa[j+1] = rev[j-1], rev[j] = q[qlen-1-j]
            const int a = j - 2, b = j - 1;
            qa[ j ] = (unsigned short)( ( ( a >= 0 && a < qlen ) ? seq.Q( qlen - 1 - a ) : 0 ) << 10 );
            qb[ j ] = (unsigned short)( ( ( b >= 0 && b < qlen ) ? seq.Q( qlen - 1 - b ) : 0 ) << 10 );
        }
    }
    int inited_end = 0;
    unsigned cells = 0; // < W * 2^16
    // H buffers: hin holds the H row of the last finished row; hbest the row of the running maximum (-1: none yet)
    int hin = 0, hbest = -1;
    int bR = 0, bSt0 = 0, bEn0 = 0;
    __half2 zmx = u2h( 0xFBFFFBFFu ); // largest unclipped maximum over the in-band cells (-65504: none yet)
    const int T0 = scM * qlen;
    unsigned char* rowBase = tb; // tb + r * ncol16
    bool stop = false;
    // r is even in every pass: the parity of c = qlen - 1 - r, hence the query copy of each of the two rows, is fixed
    const int cpar = ( qlen - 1 ) & 1;
    const unsigned* const pqA = reinterpret_cast<const unsigned*>( cpar ? sm.qb : sm.qa );
    const unsigned* const pqB = reinterpret_cast<const unsigned*>( cpar ? sm.qa : sm.qb );
    for( int r = 0; r < nrows && !stop; r += 2, rowBase += 2 * ncol16 )
    {
        const bool has2 = r + 1 < nrows;
        if( r + ( has2 ? 1 : 0 ) > w )
            return false; // the band term of the limits would become active
        // band of the two rows (the band term is inactive while r <= w)
        const int st0 = max( 0, r - qlen + 1 ), en0 = min( tlen - 1, r );
        const int st1 = max( 0, r - qlen + 2 ), en1 = has2 ? min( tlen - 1, r + 1 ) : en0 - 64 * 1024;
        const int enP = has2 ? en1 : en0; // last column of the pass
        cells += (unsigned)( en0 - st0 + 1 ) + ( has2 ? (unsigned)( en1 - st1 + 1 ) : 0u );
        if( inited_end <= enP + 1 )
        { // target codes of the columns entering the window
            const int idx = inited_end + lane;
            stc[ idx & M ] = (unsigned short)( ( idx < tlen ? seq.T( idx ) : 0 ) << 10 );
            inited_end += 32;
        }
        unsigned short fcA = fc3, fcB = fc3;
        if( r == 0 || r <= P.long_thres ) // (a negative threshold, q2 + e2 < q + e, still has the row-0 value)
        {
            fcA = r == 0 ? fc0 : r < P.long_thres ? fc1 : fc2;
            fcB = r + 1 < P.long_thres ? fc1 : r + 1 == P.long_thres ? fc2 : fc3;
        }
        if( en0 == r && lane == 0 )
            sy[ r & M ] = init6, sy2[ r & M ] = init25, su[ r & M ] = fcA;
        const int p0 = st0 & ~1;
        // left neighbour of the first pair (kswcpp_core.h:562-579), kept in the high half; column p0 - 1 is not
        // touched by row r, so row r + 1 finds the same values there (except for the first-column value)
        unsigned cX = (unsigned)init6 << 16, cX2 = (unsigned)init25 << 16, cV = (unsigned)fcA << 16,
                 cVb = (unsigned)fcB << 16;
        short* const Hin = sm.H[ hin ];
        // outputs: row r + 1 overwrites the input row in place unless that row holds the maximum
        const int o1 = hin == hbest ? ( hin + 1 ) % 3 : ( hbest < 0 ? ( hin + 1 ) % 3 : 3 - hin - hbest );
        const int o2 = hin == hbest ? ( hin + 2 ) % 3 : hin;
        unsigned* const pHin = reinterpret_cast<unsigned*>( Hin );
        unsigned* const pH1 = reinterpret_cast<unsigned*>( sm.H[ o1 ] );
        unsigned* const pH2 = reinterpret_cast<unsigned*>( sm.H[ o2 ] );
        unsigned cH = 0; // H of column p0 - 1 (high half): only read by a cell at en == p0 (never in band then)
        if( p0 > 0 )
        {
            const int kp = ( p0 - 1 ) & M;
            cX = (unsigned)sx[ kp ] << 16, cX2 = (unsigned)sx2[ kp ] << 16, cV = cVb = (unsigned)sv[ kp ] << 16;
            cH = (unsigned)(unsigned short)Hin[ kp ] << 16;
        }
        unsigned cXb = cX, cX2b = cX2, cHb = cH; // the same column after row r
        // old H left of en0, read before the pass updates it (:194-195); row 0: H[0] = v[0] - (q + e) (:247)
        const int hprev = r == 0 ? -P.qe_row0 : (int)Hin[ ( en0 - ( en0 > 0 ? 1 : 0 ) ) & M ];
        const unsigned hprev2 = ( (unsigned)hprev & 0xFFFFu ) * 0x10001u;
        __syncwarp( );
        const int c = qlen - 1 - r; // reversed-query index of column t is t + c (row r), t + c - 1 (row r + 1)
        const int qshA = c + 2 - cpar, qshB = c + cpar; // element offsets into the two copies (even)
        unsigned char* const rowpA = rowBase - ( st0 & ~15 );
        unsigned char* const rowpB = rowBase + ncol16 - ( st1 & ~15 );
        const int p1 = st1 & ~1;
        const unsigned meEnA = en0 > 0 ? 0xFFFFFFFFu : 0u;
        const unsigned injB = ( has2 && en1 == r + 1 ) ? 0xFFFFFFFFu : 0u; // column r + 1 enters in row r + 1
        const unsigned fcBx2 = (unsigned)fcB * 0x10001u;
        unsigned mA = 0x80008000u, mB = 0x80008000u, hbA = 0x80008000u, hbB = 0x80008000u;
        int t0 = p0 + 2 * lane;
        unsigned termA; // scM * (qlen - 1 - r + t) for the two cells of this lane; row r + 1: one scM less
        {
            const int a0 = scM * ( c + t0 );
            termA = ( (unsigned)a0 & 0xFFFFu ) | ( (unsigned)( a0 + scM ) << 16 );
        }
        const unsigned termStep = ( (unsigned)( scM * 64 ) & 0xFFFFu ) * 0x10001u;
        const unsigned scM2 = ( (unsigned)scM & 0xFFFFu ) * 0x10001u;
        for( int base = p0; base <= enP; base += 64, t0 += 64 )
        {
            const int kk = ( t0 & M ) >> 1; // pair index in the window
            const unsigned xo = px[ kk ], vo = pv[ kk ], x2o = px2[ kk ];
            const __half2 ut = u2h( pu[ kk ] ), yo = u2h( py[ kk ] ), y2o = u2h( py2[ kk ] );
            const unsigned hOld = pHin[ kk ];
            const __half2 tcp = u2h( ptc[ kk ] );
            const __half2 qpA = u2h( pqA[ ( ( t0 + qshA ) & M ) >> 1 ] );
            const __half2 qpB = u2h( pqB[ ( ( t0 + qshB ) & M ) >> 1 ] );
            const bool more = base + 64 <= enP;
            // ---------------- row r
            unsigned upx = __shfl_up_sync( FULL, xo, 1 ), upv = __shfl_up_sync( FULL, vo, 1 ),
                     upx2 = __shfl_up_sync( FULL, x2o, 1 );
            if( lane == 0 )
                upx = cX, upv = cV, upx2 = cX2;
            if( more )
                cX = __shfl_sync( FULL, xo, 31 ), cV = __shfl_sync( FULL, vo, 31 ), cX2 = __shfl_sync( FULL, x2o, 31 );
            unsigned z0 = sel2( __heq2_mask( tcp, qpA ), uMatch, uMis );
            z0 = sel2( __hge2_mask( __hmax2( tcp, qpA ), hCodeN ), uN, z0 );
            P2Cell A = p2_cell<LEFT>( K, u2h( __byte_perm( upx, xo, 0x5432 ) ), u2h( __byte_perm( upv, vo, 0x5432 ) ),
                                      u2h( __byte_perm( upx2, x2o, 0x5432 ) ), ut, yo, y2o, u2h( z0 ) );
            const int deA = en0 - t0; // 0: the low cell is en0, 1: the high cell
            const unsigned meA = (unsigned)deA < 2u ? ( 0xFFFFu << ( deA << 4 ) ) : 0u;
            const unsigned hA = p2_hrow( meA, meEnA, A.un, A.vn, hOld, hprev2 );
            const unsigned vmA = ( ( t0 >= st0 && deA >= 0 ) ? 0xFFFFu : 0u ) | ( deA >= 1 ? 0xFFFF0000u : 0u );
            const unsigned hmA = sel2( vmA, hA, 0x80008000u );
            mA = __vmaxs2( mA, hmA );
            hbA = __vmaxs2( hbA, __vadd2( hmA, termA & vmA ) );
            zmx = __hmax2( zmx, u2h( sel2( vmA, h2u( A.zt ), 0xFBFFFBFFu ) ) );
            if( deA >= 0 )
            {
                pH1[ kk ] = hA;
                *reinterpret_cast<unsigned short*>( rowpA + t0 ) = (unsigned short)__byte_perm( A.d, 0, 0x4420 );
            }
            // ---------------- row r + 1 on the values just computed
            const int deB = en1 - t0;
            // (with en1 == 0, a one-column target, the cell at en1 follows the interior formula: own H plus v)
            const unsigned meB = ( (unsigned)deB < 2u && en1 > 0 ) ? ( 0xFFFFu << ( deB << 4 ) ) : 0u;
            {
                const unsigned inj = meB & injB; // the entering column starts from the border values
                A.un = u2h( sel2( inj, fcBx2, h2u( A.un ) ) );
                A.yn = u2h( sel2( inj, init6x2, h2u( A.yn ) ) );
                A.y2n = u2h( sel2( inj, init25x2, h2u( A.y2n ) ) );
            }
            const unsigned xn = h2u( A.xn ), vn = h2u( A.vn ), x2n = h2u( A.x2n );
            unsigned upxb = __shfl_up_sync( FULL, xn, 1 ), upvb = __shfl_up_sync( FULL, vn, 1 ),
                     upx2b = __shfl_up_sync( FULL, x2n, 1 ), uph = __shfl_up_sync( FULL, hA, 1 );
            if( lane == 0 )
                upxb = cXb, upvb = cVb, upx2b = cX2b, uph = cHb;
            if( more )
                cXb = __shfl_sync( FULL, xn, 31 ), cVb = __shfl_sync( FULL, vn, 31 ),
                cX2b = __shfl_sync( FULL, x2n, 31 ), cHb = __shfl_sync( FULL, hA, 31 );
            unsigned z1 = sel2( __heq2_mask( tcp, qpB ), uMatch, uMis );
            z1 = sel2( __hge2_mask( __hmax2( tcp, qpB ), hCodeN ), uN, z1 );
            const P2Cell B = p2_cell<LEFT>( K, u2h( __byte_perm( upxb, xn, 0x5432 ) ),
                                            u2h( __byte_perm( upvb, vn, 0x5432 ) ),
                                            u2h( __byte_perm( upx2b, x2n, 0x5432 ) ), A.un, A.yn, A.y2n, u2h( z1 ) );
            // the cell at en1 adds u to H_r of its left neighbour
            const unsigned hB = p2_hrow( meB, 0xFFFFFFFFu, B.un, B.vn, hA, __byte_perm( uph, hA, 0x5432 ) );
            const unsigned vmB = ( ( t0 >= st1 && deB >= 0 ) ? 0xFFFFu : 0u ) | ( deB >= 1 ? 0xFFFF0000u : 0u );
            const unsigned hmB = sel2( vmB, hB, 0x80008000u );
            mB = __vmaxs2( mB, hmB );
            hbB = __vmaxs2( hbB, __vadd2( hmB, __vsub2( termA, scM2 ) & vmB ) );
            zmx = __hmax2( zmx, u2h( sel2( vmB, h2u( B.zt ), 0xFBFFFBFFu ) ) );
            termA = __vadd2( termA, termStep );
            if( has2 )
            {
                if( deB >= 0 )
                {
                    pu[ kk ] = h2u( B.un ), pv[ kk ] = h2u( B.vn ), px[ kk ] = h2u( B.xn ), py[ kk ] = h2u( B.yn );
                    px2[ kk ] = h2u( B.x2n ), py2[ kk ] = h2u( B.y2n );
                    pH2[ kk ] = hB;
                    if( t0 >= p1 )
                        *reinterpret_cast<unsigned short*>( rowpB + t0 ) = (unsigned short)__byte_perm( B.d, 0, 0x4420 );
                }
            }
        }
        const int maxA = __reduce_max_sync( FULL, max( (int)(short)( mA & 0xFFFFu ), (int)mA >> 16 ) );
        const int maxB = __reduce_max_sync( FULL, max( (int)(short)( mB & 0xFFFFu ), (int)mB >> 16 ) );
        __syncwarp( );
        int hlast = o1; // buffer of the last finished row
        // ---- row r: ksw_apply_zdrop (:22-44) with the position resolved only when it is consumed
        for( int k = 0; k < 2; k++ )
        {
            if( k == 1 && !has2 )
                break;
            const int rr = r + k, sst = k ? st1 : st0, een = k ? en1 : en0, max_H = k ? maxB : maxA, ob = k ? o2 : o1;
            hlast = ob;
            if( max_H > ez.max )
            {
                ez.max = max_H;
                hbest = ob, bR = rr, bSt0 = sst, bEn0 = een;
            }
            else if( zdrop >= 0 && ez.max - max_H > zdrop )
            {
                // resolve the position of the running maximum first
                int bt = ez.max_t, bq = ez.max_q;
                if( hbest >= 0 )
                    bt = ksw_p2_argmax( sm.H[ hbest ], M, bSt0, bEn0, lane ), bq = bR - bt;
                const int max_t = ksw_p2_argmax( sm.H[ ob ], M, sst, een, lane );
                if( max_t >= bt && rr - max_t >= bq )
                {
                    const int tl = max_t - bt, ql = ( rr - max_t ) - bq;
                    const int l = tl > ql ? tl - ql : ql - tl;
                    if( ez.max - max_H > zdrop + l * e2 )
                    {
                        ez.zdropped = 1;
                        stop = true;
                        break;
                    }
                }
            }
        }
        hin = hlast;
        if( !stop )
        {
            const int BA = __reduce_max_sync( FULL, max( (int)(short)( hbA & 0xFFFFu ), (int)hbA >> 16 ) );
            const int BB = __reduce_max_sync( FULL, max( (int)(short)( hbB & 0xFFFFu ), (int)hbB >> 16 ) );
            const int rl = r + ( has2 ? 1 : 0 ); // last row of the pass
            if( has2 && rl >= qlen )
            { // all terms are < 2^31: is16 bounds qlen, tlen and the scores
                const int j = rl + 1;
                const int T = T0 - min( q + e * j, q2 + e2 * j );
                if( max( max( BA, BB ), T ) <= ez.max )
                    stop = true;
            }
        }
    }
    if( hbest >= 0 )
    {
        __syncwarp( );
        ez.max_t = ksw_p2_argmax( sm.H[ hbest ], M, bSt0, bEn0, lane ), ez.max_q = bR - ez.max_t;
    }
    ez.cells = cells;
    __syncwarp( );
    // The no-wrap argument above needs z >= a, b, a2, b2 AND z <= match in every cell: true as long as the clip at the
    // match score (kswcpp_core.h:702) never changed a maximum. A consistent DP never clips; an inconsistent border
    // (negative long-gap threshold of swapped pieces) does, and the reference's int8 values may then wrap: such a
    // problem is redone in the banded exact mode (ksw_bx.cuh), whose arithmetic wraps like the reference's.
    return __reduce_max_sync( FULL, (int)fmaxf( __low2float( zmx ), __high2float( zmx ) ) ) <= scM;
}

#ifndef MA_KSW_MINB
#define MA_KSW_MINB 2
#endif
#ifndef MA_KSW_WARPS
#define MA_KSW_WARPS 8 // warps (DP problems in flight) per CTA
#endif
// Window classes of ksw_batch_kernel: W is the (power-of-two) window of the scalar modes and the name of the class. The
// class "1024" takes aligned band widths up to MA_KSW_CAP1024 - 48 = 592 columns — the band of 512 of the presets'
// end extensions and of the configs[4] sweep needs 544 — so that the packed banded mode runs it with a window of 640
// columns (17.5 KB of shared memory per warp: 12 warps per SM as two CTAs of six warps; a window of 1024 columns left
// 8); wider bands go to the 2048 class.
#define MA_KSW_CAP1024 640
MA_HD inline int ksw_class_cap( int W )
{
    return W == 1024 ? MA_KSW_CAP1024 : W;
}
template <int W> struct KswClass
{
    static constexpr int WB = W == 1024 ? MA_KSW_CAP1024 : W; // window of ksw_bx_rows
    static constexpr int WARPS = W == 1024 ? 6 : MA_KSW_WARPS; // warps (DP problems in flight) per CTA
    static constexpr int MINB = MA_KSW_MINB;
};

// bytes of shared memory per warp: the scalar window and, for the narrow bins, the packed one share the space
template <int W> struct KswSmemBytes
{
    static constexpr bool kPacked = W <= 512;
    static constexpr bool kBx = W <= 1024; // packed banded exact mode (ksw_bx.cuh): 28 bytes per window column and warp
    static constexpr size_t kScalar = sizeof( KswSmem<W> );
    static constexpr size_t kP2 = kPacked ? sizeof( KswSmemQ < W <= 512 ? W : 2 > ) : 0;
    static constexpr size_t kB = kBx ? sizeof( KswBxSmem < W <= 1024 ? KswClass<W>::WB : 2 > ) : 0;
    static constexpr size_t kMax2 = kScalar > kP2 ? kScalar : kP2;
    static constexpr size_t value = ( ( kMax2 > kB ? kMax2 : kB ) + 15 ) / 16 * 16;
};

template <int W>
__device__ void ksw_warp( const KswScore& P, const BxK* bxk, const bool bBx, const bool bP2, const bool bBn, const SeqAccess& seq, int qlen, int tlen, int w, int zdrop,
                          int flag, bool bEarlyStop, KswSmem<W>& sm, unsigned char* __restrict__ tb, KswOut& ez )
{
    const int lane = threadIdx.x & 31;
    ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
    ez.max = 0;
    ez.score = ez.mqe = ez.mte = (int)0x80000000;
    ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.status = 0, ez.cells = 0;
    if( qlen <= 0 || tlen <= 0 || P.early_return )
        return;
    if( w < 0 )
        w = tlen > qlen ? tlen : qlen;
    const bool bLeft = !( flag & MA_KSW_RIGHT );
    // The in-band modes are exact only while the band term of the limits is inactive (rows r <= w) and hand over to
    // the exact mode otherwise, which repeats the work. The early-stop bound cannot fire before the last query row
    // has been reached on the main diagonal (r ~ 2 qlen), so they are only tried where that row lies inside r <= w
    // (or the whole matrix does): the alignment path's end extensions (w = 512, read tails) always qualify.
    const bool bInBandPays = w >= qlen && ( 2 * qlen + 16 <= w || qlen + tlen - 1 <= w + 1 );
    if constexpr( KswSmemBytes<W>::kPacked )
    { // packed fast path: early-stop extensions in int16 score mode whose int8 arithmetic cannot wrap
        const int iSize = qlen > tlen ? qlen : tlen;
        const bool is16 = !( (long long)iSize * P.min16 < -32768 || (long long)iSize * P.match > 32767 );
        if( bP2 && bEarlyStop && bInBandPays && is16 && qlen + 4 <= W && ( qlen < tlen ? qlen : tlen ) + 40 <= W &&
            ksw_p2_params_ok( P ) )
        {
            KswSmemQ<W>& sp = reinterpret_cast<KswSmemQ<W>&>( sm );
            const bool ok = bLeft ? ksw_rows_p2x2<W, true>( P, seq, qlen, tlen, w, zdrop, sp, tb, ez )
                                  : ksw_rows_p2x2<W, false>( P, seq, qlen, tlen, w, zdrop, sp, tb, ez );
            if( ok )
                return;
            ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
            ez.max = 0;
            ez.score = ez.mqe = ez.mte = (int)0x80000000;
            ez.zdropped = 0, ez.cells = 0;
            __syncwarp( );
        }
    }
    if constexpr( KswSmemBytes<W>::kBx )
    { // packed banded exact mode: scoring parameters whose int8 arithmetic cannot wrap
        if( bBx && ksw_bx_params_ok( P ) )
        {
            if constexpr( W <= 256 )
            { // register-resident narrow-band kernel (ksw_bn.cuh): aligned ranges of at most three 64-column chunks
                static_assert( sizeof( KswBnSmem ) <= KswSmemBytes<W>::value, "narrow-band scratch" );
                const int nch = ( ksw_ncol16( qlen, tlen, w ) + 63 ) / 64;
                if( bBn && nch <= MA_BN_MAXC )
                {
                    KswBnSmem& sn = reinterpret_cast<KswBnSmem&>( sm );
#define MA_BN_CALL( N )                                                                                                \
    if( bLeft )                                                                                                        \
        ksw_bn_rows<N, true>( bxk[ 0 ], P, seq, qlen, tlen, w, zdrop, bEarlyStop, sn, tb, ez );                        \
    else                                                                                                               \
        ksw_bn_rows<N, false>( bxk[ 1 ], P, seq, qlen, tlen, w, zdrop, bEarlyStop, sn, tb, ez );
                    if( W == 128 && nch == 1 )
                    {
                        MA_BN_CALL( 1 )
                    }
                    else if( nch <= 2 )
                    {
                        MA_BN_CALL( 2 )
                    }
                    else
                    {
                        MA_BN_CALL( 3 )
                    }
#undef MA_BN_CALL
                    return;
                }
            }
            constexpr int WB = KswClass<W>::WB;
            KswBxSmem<WB>& sb = reinterpret_cast<KswBxSmem<WB>&>( sm );
            if( bLeft )
                ksw_bx_rows<WB, true>( bxk[ 0 ], P, seq, qlen, tlen, w, zdrop, bEarlyStop, sb, tb, ez );
            else
                ksw_bx_rows<WB, false>( bxk[ 1 ], P, seq, qlen, tlen, w, zdrop, bEarlyStop, sb, tb, ez );
            return;
        }
    }
    // stage the query in shared memory when it fits
    const bool qStaged = qlen <= KswSmem<W>::QC;
    if( qStaged )
    {
        for( int i = lane; i < qlen; i += 32 )
            sm.qc[ i ] = (unsigned char)seq.Q( i );
        __syncwarp( );
    }
    // FAST mode needs the query staged (narrow problems) and a band that cannot limit before the matrix does
    if( bEarlyStop && qStaged && bInBandPays )
    {
        const bool ok = bLeft ? ksw_rows<W, true, true, true>( P, seq, qlen, tlen, w, zdrop, true, sm, tb, ez )
                              : ksw_rows<W, false, true, true>( P, seq, qlen, tlen, w, zdrop, true, sm, tb, ez );
        if( ok )
            return;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
        ez.max = 0;
        ez.score = ez.mqe = ez.mte = (int)0x80000000;
        ez.zdropped = 0, ez.cells = 0;
        __syncwarp( );
    }
    if( qStaged )
    {
        if( bLeft )
            ksw_rows<W, true, true, false>( P, seq, qlen, tlen, w, zdrop, bEarlyStop, sm, tb, ez );
        else
            ksw_rows<W, false, true, false>( P, seq, qlen, tlen, w, zdrop, bEarlyStop, sm, tb, ez );
    }
    else
    {
        if( bLeft )
            ksw_rows<W, true, false, false>( P, seq, qlen, tlen, w, zdrop, bEarlyStop, sm, tb, ez );
        else
            ksw_rows<W, false, false, false>( P, seq, qlen, tlen, w, zdrop, bEarlyStop, sm, tb, ez );
    }
}

struct KswBatchArgs
{
    const KswTask* tasks;
    const int* order; // task indices of this launch (binned by window size)
    int n;
    const unsigned char* seq; // byte slab (standalone batches) or the read slab (pipeline)
    const unsigned char* pac; // pack of the uploaded index (pipeline tasks with MA_TASK_TPACK)
    long long fwd_len;
    KswOut* out;
    unsigned int* cigar; // output slab
    long long cigar_cap; // words
    unsigned long long* cigar_cursor;
    unsigned char* tb; // per-warp traceback slabs
    long long tb_stride; // bytes per warp
    unsigned int* cigscratch; // per-warp cigar scratch
    int cigscratch_stride; // words per warp
    int* next; // dynamic task queue
    int* error;
    unsigned long long* cells_total; // optional: sum of band cells (GCUPS accounting)
    KswScore score;
    // ksw_qs_kernel only: where a problem that left the kernel's regime is handed over to ksw_batch_kernel.
    // Pipeline: the device-side bins of nwbin_kernel (redo_count != nullptr); standalone batches: one list.
    int* redo_count; // [15] tasks per (window class, kind) bin
    unsigned long long* redo_tb; // [15] traceback bytes per warp
    int* redo_cig; // [15] cigar scratch words per warp
    int* redo_order; // [bins][redo_cap] task ids (pipeline) / [redo_cap] (list form, its length in redo_n)
    long long redo_cap;
    int* redo_n;
    QsK qsk[ 2 ]; // packed constants of ksw_qs_kernel: [0] left-aligned, [1] right-aligned
    BxK bxk[ 2 ]; // ... of ksw_bx_rows
    int no_bx; // bit 0: scalar exact mode instead of ksw_bx_rows, bit 1: no ksw_rows_p2x2, bit 2: no ksw_bn_rows (measurements)
};

// window class of ksw_batch_kernel for an aligned band width
MA_HD inline int ksw_bin_of( int ncol16 )
{
    const int need = ncol16 + 48;
    return need <= 128 ? 0 : need <= 256 ? 1 : need <= 512 ? 2 : need <= MA_KSW_CAP1024 ? 3 : need <= 2048 ? 4 : 5;
}

template <int W>
__global__ void __launch_bounds__( 32 * KswClass<W>::WARPS, KswClass<W>::MINB ) ksw_batch_kernel( KswBatchArgs A )
{
    extern __shared__ __align__( 16 ) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    KswSmem<W>& sm = *reinterpret_cast<KswSmem<W>*>( smem_raw + (size_t)warp * KswSmemBytes<W>::value );
    const long long gw = (long long)blockIdx.x * ( blockDim.x >> 5 ) + warp;
    unsigned char* tb = A.tb + gw * A.tb_stride;
    unsigned int* cs = A.cigscratch + gw * A.cigscratch_stride;
    while( true )
    {
        int slot = 0;
        if( lane == 0 )
            slot = atomicAdd( A.next, 1 );
        slot = __shfl_sync( 0xffffffffu, slot, 0 );
        if( slot >= A.n )
            break;
        const int ti = A.order[ slot ];
        const KswTask T = A.tasks[ ti ];
        KswOut ez;
        SeqAccess sa;
        sa.qbase = A.seq, sa.qoff = T.qoff, sa.qstep = ( T.tag & MA_TASK_QREV ) ? -1 : 1;
        sa.tslab = A.seq, sa.toff = T.toff, sa.tstep = ( T.tag & MA_TASK_TREV ) ? -1 : 1;
        sa.pac = ( T.tag & MA_TASK_TPACK ) ? A.pac : nullptr, sa.fwd_len = A.fwd_len;
        ksw_warp<W>( A.score, A.bxk, ( A.no_bx & 1 ) == 0, ( A.no_bx & 2 ) == 0, ( A.no_bx & 4 ) == 0, sa, T.qlen, T.tlen, T.w, T.zdrop, T.flag, ( T.tag & MA_TASK_EARLYSTOP ) != 0, sm, tb, ez );
        ez.cigar_off = 0;
        int i0 = 0, j0 = 0, n = 0;
        const bool bBt = ( T.qlen > 0 && T.tlen > 0 && !A.score.early_return ) &&
                         ksw_bt_start( ez, T.qlen, T.tlen, T.flag, i0, j0 );
        if( bBt )
        {
            if( lane == 0 )
                n = ksw_backtrack( tb, ksw_ncol16( T.qlen, T.tlen, T.w ), T.qlen, T.tlen,
                                   T.w < 0 ? ( T.tlen > T.qlen ? T.tlen : T.qlen ) : T.w, i0, j0, cs,
                                   A.cigscratch_stride );
            n = __shfl_sync( 0xffffffffu, n, 0 );
            unsigned long long o = 0;
            if( n > 0 && lane == 0 )
                o = atomicAdd( A.cigar_cursor, (unsigned long long)n );
            o = __shfl_sync( 0xffffffffu, o, 0 );
            if( n < 0 || (long long)( o + ( n > 0 ? n : 0 ) ) > A.cigar_cap )
            {
                if( lane == 0 )
                    atomicExch( A.error, 1 );
                ez.status = 1;
                n = 0;
            }
            __syncwarp( );
            // the walk produced ops end-to-start; the reference reverses unless REV_CIGAR is set
            const bool rev = T.flag & MA_KSW_REV_CIGAR;
            for( int k = lane; k < n; k += 32 )
                A.cigar[ o + k ] = rev ? cs[ k ] : cs[ n - 1 - k ];
            ez.n_cigar = n;
            ez.cigar_off = (long long)o;
        }
        if( lane == 0 )
        {
            A.out[ ti ] = ez;
            if( A.cells_total )
                atomicAdd( A.cells_total, (unsigned long long)ez.cells );
        }
        __syncwarp( );
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Query-stationary register kernel (ksw_qs.cuh): one warp per early-stop extension with a query of <= 64 NB bases.
#ifndef MA_QS_WARPS
#define MA_QS_WARPS 8
#endif
#ifndef MA_QS_MINB
#define MA_QS_MINB 3
#endif
// bytes of traceback one warp needs for a problem of this kernel
MA_HD inline long long ksw_qs_tb_bytes( int nb, int qlen, int tlen, int w )
{
    if( w < 0 )
        w = tlen > qlen ? tlen : qlen;
    const long long rows = (long long)qlen + tlen < (long long)w + 2 ? (long long)qlen + tlen : (long long)w + 2;
    return ( ( rows + 1 ) * 64 * nb + 255 ) & ~255ll; // (rows are stored in pairs)
}

template <int NB, bool LEFT>
__global__ void __launch_bounds__( 32 * MA_QS_WARPS, MA_QS_MINB ) ksw_qs_kernel( KswBatchArgs A )
{
    __shared__ KswQsSmem<NB> smAll[ MA_QS_WARPS ];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    KswQsSmem<NB>& sm = smAll[ warp ];
    const long long gw = (long long)blockIdx.x * MA_QS_WARPS + warp;
    unsigned char* tb = A.tb + gw * A.tb_stride;
    unsigned int* cs = A.cigscratch + gw * A.cigscratch_stride;
    unsigned long long cellsLocal = 0;
    while( true )
    {
        int slot = 0;
        if( lane == 0 )
            slot = atomicAdd( A.next, 1 );
        slot = __shfl_sync( 0xffffffffu, slot, 0 );
        if( slot >= A.n )
            break;
        const int ti = A.order[ slot ];
        const KswTask T = A.tasks[ ti ];
        KswOut ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
        ez.max = 0;
        ez.score = ez.mqe = ez.mte = (int)0x80000000;
        ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.status = 0, ez.cells = 0, ez.cigar_off = 0;
        SeqAccess sa;
        sa.qbase = A.seq, sa.qoff = T.qoff, sa.qstep = ( T.tag & MA_TASK_QREV ) ? -1 : 1;
        sa.tslab = A.seq, sa.toff = T.toff, sa.tstep = ( T.tag & MA_TASK_TREV ) ? -1 : 1;
        sa.pac = ( T.tag & MA_TASK_TPACK ) ? A.pac : nullptr, sa.fwd_len = A.fwd_len;
        const int w = T.w < 0 ? ( T.tlen > T.qlen ? T.tlen : T.qlen ) : T.w;
        const bool ok = ksw_qs_rows<NB, LEFT>( A.qsk[ LEFT ? 0 : 1 ], A.score, sa, T.qlen, T.tlen, w, T.zdrop, sm, tb, ez );
        if( !ok )
        { // hand the problem over to ksw_batch_kernel
            if( lane == 0 )
            {
                if( A.redo_count )
                {
                    const int nc = ksw_ncol16( T.qlen, T.tlen, T.w );
                    const int wc = ksw_bin_of( nc );
                    const int b = wc < 5 ? wc * 3 + ( LEFT ? 1 : 2 ) : 15;
                    const int s = atomicAdd( &A.redo_count[ b ], 1 );
                    if( s < A.redo_cap )
                        A.redo_order[ (long long)b * A.redo_cap + s ] = ti;
                    atomicMax( &A.redo_tb[ b ], ( ( (unsigned long long)T.qlen + T.tlen ) * nc + 255 ) & ~255ull );
                    atomicMax( &A.redo_cig[ b ], ( T.qlen + T.tlen + 2 + 63 ) & ~63 );
                }
                else
                {
                    const int s = atomicAdd( A.redo_n, 1 );
                    if( s < A.redo_cap )
                        A.redo_order[ s ] = ti;
                }
            }
            __syncwarp( );
            continue;
        }
        int n = 0;
        if( ez.max_t >= 0 && ez.max_q >= 0 )
        {
            __syncwarp( );
            if( lane == 0 )
                n = ksw_qs_backtrack( tb, 64 * NB, LEFT, T.qlen, T.tlen, w, ez.max_t, ez.max_q, cs, A.cigscratch_stride );
            n = __shfl_sync( 0xffffffffu, n, 0 );
            unsigned long long o = 0;
            if( n > 0 && lane == 0 )
                o = atomicAdd( A.cigar_cursor, (unsigned long long)n );
            o = __shfl_sync( 0xffffffffu, o, 0 );
            if( n < 0 || (long long)( o + ( n > 0 ? n : 0 ) ) > A.cigar_cap )
            {
                if( lane == 0 )
                    atomicExch( A.error, 1 );
                ez.status = 1;
                n = 0;
            }
            __syncwarp( );
            const bool rev = T.flag & MA_KSW_REV_CIGAR;
            for( int k = lane; k < n; k += 32 )
                A.cigar[ o + k ] = rev ? cs[ k ] : cs[ n - 1 - k ];
            ez.n_cigar = n;
            ez.cigar_off = (long long)o;
        }
        if( lane == 0 )
            A.out[ ti ] = ez;
        cellsLocal += (unsigned long long)ez.cells;
        __syncwarp( );
    }
    if( lane == 0 && A.cells_total && cellsLocal )
        atomicAdd( A.cells_total, cellsLocal );
}

} // namespace ma
