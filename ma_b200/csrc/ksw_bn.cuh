// Banded exact DP for NARROW bands, register resident (sm_100a, DPX): the arithmetic and the reference semantics of
// ksw_bx.cuh (kswcpp_inner_core with its 16-aligned column ranges, stale out-of-band cells, int8 wrap-around, every
// kswcpp_extz_t field), for problems whose aligned range never exceeds 64 NC columns (NC <= 3: bands up to 128).
//
// Why. With a band of 16 .. 128 cells a row is one to three 64-column chunks, and the shared-memory window of
// ksw_bx.cuh pays ~450 instructions of per-row bookkeeping, two __syncwarp and a load / store of all state for 17 .. 129
// useful cells (configs[4]: 30 - 70 GCUPS, 61 % of the sweep's time). Here a lane OWNS the column pairs
// (base + 64 c + 2 lane, + 1), c < NC, of a window that starts at the aligned band start `base`, and keeps their state —
// u, v, x, y, x2, y2, the stale score profile, the target codes and H — in REGISTERS over the whole problem:
//   * the left neighbour is one shuffle per array, the H of en0's left neighbour and the row's H[en0] / H[st0] are
//     single shuffles: a row touches shared memory once (the query pair) and has no barrier;
//   * when the aligned start moves on by 16 columns (every ~32 rows) all arrays shift down by 8 lanes (one shuffle each)
//     and the 16 columns that enter start from the reference's initial values;
//   * the position of the row maximum is resolved only when it is consumed, from 1 KB copies of the H row in shared
//     memory (the row of the running maximum is written there when it becomes the maximum; ksw_bx_argmax).
// Traceback bytes, CIGAR walk, early termination: as in ksw_bx.cuh.
#pragma once
#include "ksw_bx.cuh"

namespace ma
{

struct KswBnSmem
{
    unsigned QE[ 128 ], QO[ 128 ]; // query window of 256 bases in two copies (ksw_bx.cuh)
    int HS[ 256 ], HBS[ 256 ]; // scratch rows for ksw_bx_argmax: H of the current row / of the row of the running maximum
};

// largest number of 64-column chunks of this kernel
#define MA_BN_MAXC 3

template <int NC, bool LEFT>
QS_DEV void ksw_bn_rows( const BxK& K, const KswScore& P, const SeqAccess& seq, const int qlen, const int tlen, const int w,
                         const int zdrop, const bool bEarlyStop, KswBnSmem& sm, unsigned char* __restrict__ tb, KswOut& ez )
{
    const unsigned FULL = 0xffffffffu;
    const int lane = qs_lane( );
    const int MPQ = 127;
    const int NEG_M = -0x40000000; // removes a cell from a maximum
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, scM = P.match;
    const int T16 = ( ( tlen + 15 ) / 16 ) * 16;
    const int ncol16 = ksw_ncol16( qlen, tlen, w );
    const int iSize = qlen > tlen ? qlen : tlen;
    const bool is16 = !( (long long)iSize * P.min16 < -32768 || (long long)iSize * P.match > 32767 );
    const int SMASK = is16 ? ~7 : ~3; // SSE lanes of the reference's H vectors: 8 x int16 or 4 x int32
    const int NEG_INF = is16 ? -32768 : (int)0x80000000;
    const int nrows = qlen + tlen - 1;
    // state of the column pairs of this lane
    unsigned U[ NC ], V[ NC ], X[ NC ], Y[ NC ], X2[ NC ], Y2[ NC ], S[ NC ], TC[ NC ];
    int H0[ NC ], H1[ NC ]; // H (the H row of the running maximum is kept in sm.HBS)
    bool anyN = false;
    auto tcodes = [ & ]( const int t0 ) {
        const int c0 = t0 < tlen ? seq.T( t0 ) : 0, c1 = t0 + 1 < tlen ? seq.T( t0 + 1 ) : 0;
        anyN |= c0 >= 4 || c1 >= 4;
        return ( (unsigned)c0 << 10 ) | ( (unsigned)c1 << 26 );
    };
#pragma unroll
    for( int c = 0; c < NC; c++ )
    {
        U[ c ] = V[ c ] = K.iUV, X[ c ] = K.iX, Y[ c ] = K.iY, X2[ c ] = K.iX2, Y2[ c ] = K.iY2, S[ c ] = K.iS;
        H0[ c ] = H1[ c ] = NEG_INF;
        TC[ c ] = tcodes( 64 * c + 2 * lane );
    }
    anyN = __any_sync( FULL, anyN );
    int base = 0; // first column of the register window: the aligned band start
    int qend = -32; // query bases j < qend are staged
    unsigned cells = 0;
    int prevB = 0x7fffffff;
    int bR = -1, bSt0 = 0, bEn0 = 0, bT = -1; // row, band and (once resolved) position of the running maximum
    int Hleft = NEG_INF; // H of column base - 1 (it left the window; read once more when the band is the one column `base`)
    // value of column t of a two-register row, on all lanes
    auto colval = [ & ]( const int( &A0 )[ NC ], const int( &A1 )[ NC ], const int t, const int wbase ) {
        const int d = t - wbase;
        int v = 0; // (masks instead of branches: the arrays must stay in registers)
#pragma unroll
        for( int c = 0; c < NC; c++ )
            v |= ( ( d & 1 ) ? A1[ c ] : A0[ c ] ) & ( ( d >> 6 ) == c ? -1 : 0 );
        return __shfl_sync( FULL, v, ( d >> 1 ) & 31 );
    };
    // copies the H row into a scratch row (column t at [t & 255])
    auto spill = [ & ]( int* scratch ) {
        __syncwarp( );
#pragma unroll
        for( int c = 0; c < NC; c++ )
        {
            const int t0 = base + 64 * c + 2 * lane;
            scratch[ t0 & 255 ] = H0[ c ], scratch[ ( t0 & 255 ) + 1 ] = H1[ c ];
        }
        __syncwarp( );
    };
    unsigned char* rowp0 = tb; // tb + r * ncol16
    for( int r = 0; r < nrows; ++r, rowp0 += ncol16 )
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
        cells += (unsigned)( en0 - st0 + 1 );
        const int st = st0 & ~15, en = en0 | 15;
        const int sEnd = bx_min( st0 + ( ( ( en0 - st0 ) >> 4 ) + 1 ) * 16, T16 ); // score-profile end
        // values entering the first column from its left neighbour (:562-579), kept in the HIGH half
        unsigned cX = K.iX & 0xFFFF0000u, cX2 = K.iX2 & 0xFFFF0000u, cV = K.iUV & 0xFFFF0000u;
        if( st > base )
        { // the window moves on by 16 columns; column st - 1 (computed in the last row: lane 7, high half) is the carry
            cX = __shfl_sync( FULL, X[ 0 ], 7 ) & 0xFFFF0000u, cX2 = __shfl_sync( FULL, X2[ 0 ], 7 ) & 0xFFFF0000u;
            cV = __shfl_sync( FULL, V[ 0 ], 7 ) & 0xFFFF0000u;
            Hleft = __shfl_sync( FULL, H1[ 0 ], 7 );
            const int src = ( lane + 8 ) & 31;
            const bool low = lane < 24;
            base = st;
            const unsigned tcNew = low ? 0u : tcodes( base + 64 * ( NC - 1 ) + 2 * lane );
            anyN = __any_sync( FULL, anyN );
#pragma unroll
            for( int c = 0; c < NC; c++ )
            {
                const bool last = c + 1 == NC;
#define MA_BN_SHIFT( A, INIT )                                                                                         \
    {                                                                                                                  \
        const auto a_ = __shfl_sync( FULL, A[ c ], src );                                                              \
        const auto b_ = last ? ( INIT ) : __shfl_sync( FULL, A[ last ? c : c + 1 ], src );                             \
        A[ c ] = low ? a_ : b_;                                                                                        \
    }
                MA_BN_SHIFT( U, K.iUV )
                MA_BN_SHIFT( V, K.iUV )
                MA_BN_SHIFT( X, K.iX )
                MA_BN_SHIFT( Y, K.iY )
                MA_BN_SHIFT( X2, K.iX2 )
                MA_BN_SHIFT( Y2, K.iY2 )
                MA_BN_SHIFT( S, K.iS )
                MA_BN_SHIFT( TC, tcNew )
                MA_BN_SHIFT( H0, NEG_INF )
                MA_BN_SHIFT( H1, NEG_INF )
#undef MA_BN_SHIFT
            }
        }
        else if( st == 0 )
        {
            const int first_col = r == 0 ? -q - e : r < P.long_thres ? -e : r == P.long_thres ? P.long_diff : -e2;
            cV = (unsigned)( first_col * 256 ) << 16;
        }
        while( qend <= r - st0 + 1 )
        { // the next 32 query bases: lanes 0-15 write QE, lanes 16-31 QO
            __syncwarp( ); // (the entries they replace were read 128+ rows ago; the rows of this kernel have no barrier)
            const int k = ( qend >> 1 ) + ( lane & 15 );
            const int jl = 2 * k + ( lane >> 4 );
            const int c0 = (unsigned)jl < (unsigned)qlen ? seq.Q( jl ) : 0;
            const int c1 = (unsigned)( jl - 1 ) < (unsigned)qlen ? seq.Q( jl - 1 ) : 0;
            ( lane < 16 ? sm.QE : sm.QO )[ k & MPQ ] = ( (unsigned)c0 << 10 ) | ( (unsigned)c1 << 26 );
            anyN |= __any_sync( FULL, c0 >= 4 || c1 >= 4 );
            qend += 32;
            __syncwarp( );
        }
        if( en >= r )
        { // column r enters through the first query row: y, y2, u start from the border values
            const int first_col = r == 0 ? -q - e : r < P.long_thres ? -e : r == P.long_thres ? P.long_diff : -e2;
            const int d = r - base;
            const unsigned hm = ( d & 1 ) ? 0xFFFF0000u : 0x0000FFFFu;
            const unsigned fc = ( (unsigned)( first_col * 256 ) & 0xFFFFu ) * 0x10001u;
#pragma unroll
            for( int c = 0; c < NC; c++ )
            { // (masks instead of branches: the arrays must stay in registers)
                const unsigned hmc = ( ( d >> 6 ) == c && ( ( d >> 1 ) & 31 ) == lane ) ? hm : 0u;
                Y[ c ] = ( Y[ c ] & ~hmc ) | ( K.iY & hmc ), Y2[ c ] = ( Y2[ c ] & ~hmc ) | ( K.iY2 & hmc );
                U[ c ] = ( U[ c ] & ~hmc ) | ( fc & hmc );
            }
        }
        // old H left of en0, read before the pass updates it (:194-195); row 0: H[0] = v[0] - (q + e) (:247)
        int hprev = -P.qe_row0;
        if( r > 0 )
            hprev = ( en0 > 0 && en0 - 1 < base ) ? Hleft : colval( H0, H1, en0 > 0 ? en0 - 1 : en0, base );
        const unsigned nS = (unsigned)( sEnd - st0 ), nB = (unsigned)( en0 - st0 );
        const unsigned* const qw = ( r & 1 ) ? sm.QO : sm.QE; // r - t0 has the parity of r
        int m = NEG_M, hb = NEG_M;
        unsigned char* const rowp = rowp0 - st;
        // chunks in descending order: chunk c reads the old values of chunk c - 1 (its left neighbour at lane 0)
#pragma unroll
        for( int c = NC - 1; c >= 0; c-- )
        {
            unsigned upx = __shfl_up_sync( FULL, X[ c ], 1 ), upv = __shfl_up_sync( FULL, V[ c ], 1 ),
                     upx2 = __shfl_up_sync( FULL, X2[ c ], 1 );
            if( c > 0 )
            {
                const unsigned px = __shfl_sync( FULL, X[ c > 0 ? c - 1 : 0 ], 31 ), pv = __shfl_sync( FULL, V[ c > 0 ? c - 1 : 0 ], 31 ),
                               px2 = __shfl_sync( FULL, X2[ c > 0 ? c - 1 : 0 ], 31 );
                if( lane == 0 )
                    upx = px, upv = pv, upx2 = px2;
            }
            else if( lane == 0 )
                upx = cX, upv = cV, upx2 = cX2;
            const int t0 = base + 64 * c + 2 * lane;
            const unsigned qp = qw[ ( ( r - t0 ) >> 1 ) & MPQ ];
            // score profile (:591-616); cells outside [st0, sEnd) keep the stale profile of their column
            unsigned z0 = ( qs_eqmask2( TC[ c ], qp ) & K.zXor ) ^ K.zMis;
            if( anyN )
            {
                const unsigned nm = bx_nmask2( TC[ c ], qp );
                z0 = ( K.zN & nm ) | ( z0 & ~nm );
            }
            const unsigned d0 = (unsigned)( t0 - st0 ), d1 = d0 + 1u;
            const unsigned fm = ( d0 < nS ? 0xFFFFu : 0u ) | ( d1 < nS ? 0xFFFF0000u : 0u );
            z0 = ( z0 & fm ) | ( S[ c ] & ~fm );
            S[ c ] = z0; // (also the profile entries the reference writes beyond the aligned range)
            const BxCell C = ksw_bx_cell<LEFT>( K, __byte_perm( upx, X[ c ], 0x5432 ), __byte_perm( upv, V[ c ], 0x5432 ),
                                                __byte_perm( upx2, X2[ c ], 0x5432 ), U[ c ], Y[ c ], Y2[ c ], z0 );
            if( t0 <= en && t0 >= st )
            { // a column of the aligned range
                // H row: interior columns add v to their own H, the column at en0 adds u to the old H of its left
                // neighbour (v in column 0); H of a column left of the band stays
                int h0 = (int)( (unsigned)H0[ c ] + (unsigned)( (int)( C.vn << 16 ) >> 24 ) );
                int h1 = (int)( (unsigned)H1[ c ] + (unsigned)( (int)C.vn >> 24 ) );
                if( t0 == en0 )
                    h0 = (int)( (unsigned)hprev + (unsigned)( (int)( ( en0 > 0 ? C.un : C.vn ) << 16 ) >> 24 ) );
                if( t0 + 1 == en0 )
                    h1 = (int)( (unsigned)hprev + (unsigned)( (int)C.un >> 24 ) );
                if( is16 )
                    h0 = (short)h0, h1 = (short)h1;
                const bool in0 = d0 <= nB, in1 = d1 <= nB; // t in [st0, en0]
                if( t0 >= st0 )
                    H0[ c ] = h0;
                if( t0 + 1 >= st0 )
                    H1[ c ] = h1;
                const int hm0 = in0 ? h0 : NEG_M, hm1 = in1 ? h1 : NEG_M;
                m = bx_max( m, bx_max( hm0, hm1 ) );
                if( bEarlyStop )
                { // scM * (query rows still below the cell), see ksw.cuh
                    const int term0 = scM * ( qlen - 1 - r + t0 );
                    hb = bx_max( hb, bx_max( hm0 + term0, hm1 + term0 + scM ) );
                }
                U[ c ] = C.un, V[ c ] = C.vn, X[ c ] = C.xn, Y[ c ] = C.yn, X2[ c ] = C.x2n, Y2[ c ] = C.y2n;
                *reinterpret_cast<unsigned short*>( rowp + t0 ) = (unsigned short)__byte_perm( C.tbyte, 0, 0x4420 );
            }
        }
        // the row maximum is the exact maximum of the row (only its POSITION is lane-blocked in the reference)
        const int max_H = __reduce_max_sync( FULL, m );
        int scoreLast = 0;
        if( en0 == tlen - 1 )
        { // (H[en0] is only consumed in the last target column: mte, and score in the last row)
            const int Hen0 = colval( H0, H1, en0, base );
            if( Hen0 > ez.mte )
                ez.mte = Hen0, ez.mte_q = r - en; // sic: the aligned en
            scoreLast = Hen0;
        }
        if( r - st0 == qlen - 1 )
        {
            const int Hst0 = colval( H0, H1, st0, base );
            if( Hst0 > ez.mqe )
                ez.mqe = Hst0, ez.mqe_t = st0;
        }
        // ksw_apply_zdrop (:22-44); the position of a new maximum is resolved when a later test or the end consumes it
        if( max_H > ez.max )
        {
            ez.max = max_H;
            bR = r, bSt0 = st0, bEn0 = en0, bT = -1;
            spill( sm.HBS );
        }
        else if( zdrop >= 0 && ez.max - max_H > zdrop )
        {
            if( bR >= 0 && bT < 0 ) // resolved once per maximum
                bT = ksw_bx_argmax<256>( sm.HBS, bSt0, bEn0, lane, SMASK );
            const int bt = bT, bq = bR >= 0 ? bR - bT : -1;
            spill( sm.HS );
            const int max_t = ksw_bx_argmax<256>( sm.HS, st0, en0, lane, SMASK );
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
        if( r == nrows - 1 && en0 == tlen - 1 ) // (after the z-drop test: a drop in the last row leaves no score)
            ez.score = scoreLast;
        if( bEarlyStop )
        { // see ksw.cuh, ksw_rows
            const int B = __reduce_max_sync( FULL, hb );
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
        ez.max_t = bT >= 0 ? bT : ksw_bx_argmax<256>( sm.HBS, bSt0, bEn0, lane, SMASK ), ez.max_q = bR - ez.max_t;
    ez.cells = cells;
    __syncwarp( );
}

} // namespace ma
