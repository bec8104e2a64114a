// CPU ORACLE — test infrastructure only. Never linked into, called by, or shipped with the product library.
//
// Scalar restatement of the reference's banded two-piece-affine DP (kswcpp, a re-templating of ksw2 extd2):
//   /root/reference/libs/kswcpp/inc/kswcpp_core.h:308-841   kswcpp_inner_core
//   /root/reference/libs/kswcpp/inc/kswcpp_core.h:157-299   calcMaxScore (exact-max branch only; MA never sets APPROX_MAX)
//   /root/reference/libs/kswcpp/inc/kswcpp_core.h:22-44     ksw_apply_zdrop
//   /root/reference/libs/kswcpp/inc/kswcpp_core.h:76-150    ksw_backtrack__
//   /root/reference/libs/kswcpp/src/kswcpp_sse_xx.cpp:38-68 int16 / int32 score width switch
// The reference computes with 16 x int8 SSE vectors on 16-ALIGNED column ranges; cells outside the band but inside
// the aligned range are computed from stale state and DO feed band-edge cells, so this restatement keeps the exact
// same arrays (u,v,x,y,x2,y2,s over the target index), the same aligned ranges and the same write patterns.
// Parity is pinned against the compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py.
#include "oracle.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <vector>

namespace
{
inline int8_t w8( int x ) // int8 wrap-around, the behaviour of _mm_add_epi8 / _mm_sub_epi8
{
    return (int8_t)(uint8_t)x;
}

struct Cigar
{
    std::vector<uint32_t> v;
    void push( uint32_t op, int len ) // kswcpp_core.h:46-65
    {
        if( v.empty( ) || op != ( v.back( ) & 0xf ) )
            v.push_back( (uint32_t)len << 4 | op );
        else
            v.back( ) += (uint32_t)len << 4;
    }
};

// kswcpp_core.h:76-150 (is_rot = 1, min_intron_len = 0)
void backtrack( bool is_rev, const std::vector<uint8_t>& p, const std::vector<int>& off, const std::vector<int>& off_end,
                int64_t n_col, int i0, int j0, Cigar& cig )
{
    int64_t i = i0, j = j0, r;
    int state = 0;
    while( i >= 0 && j >= 0 )
    {
        int force_state = -1;
        r = i + j;
        if( i < off[ r ] )
            force_state = 2;
        if( i > off_end[ r ] )
            force_state = 1;
        uint32_t tmp = force_state < 0 ? p[ r * n_col + i - off[ r ] ] : 0;
        if( state == 0 )
            state = tmp & 7;
        else if( !( tmp >> ( state + 2 ) & 1 ) )
            state = 0;
        if( state == 0 )
            state = tmp & 7;
        if( force_state >= 0 )
            state = force_state;
        if( state == 0 )
            cig.push( 0, 1 ), --i, --j;
        else if( state == 1 || state == 3 )
            cig.push( 2, 1 ), --i;
        else
            cig.push( 1, 1 ), --j;
    }
    if( i >= 0 )
        cig.push( 2, (int)i + 1 );
    if( j >= 0 )
        cig.push( 1, (int)j + 1 );
    if( !is_rev )
        std::reverse( cig.v.begin( ), cig.v.end( ) );
}

// Instrumentation for validating the product's early-termination bound: per row
//   B_r = max over in-band cells of H + match * (qlen - 1 - i),  T_r = match * qlen - mingap(r + 1);
// the product stops an EXTZ_ONLY problem after row r >= qlen when max(B_r, B_{r-1}, T_r) <= ez.max.
// The checker records the first such row and whether ANY later row still raised ez.max (a violation).
struct EarlyStopCheck
{
    int64_t stop_row = -1, rows = 0, violated = 0;
};
static thread_local EarlyStopCheck* g_pEsc = nullptr;

template <typename TS, int SIZE>
void core( int qlen, const uint8_t* query, int tlen, const uint8_t* target, const ma_oracle_score_t& sc, int w,
           int zdrop, int flag, ma_oracle_ksw_t* ez, Cigar& cig, int64_t* pCells )
{
    int64_t escPrevB = std::numeric_limits<int64_t>::max( );
    const TS NEG_INF = std::numeric_limits<TS>::min( );
    int8_t q = (int8_t)sc.gap, e = (int8_t)sc.extend, q2 = (int8_t)sc.gap2, e2 = (int8_t)sc.extend2;
    const int8_t sc_mch = (int8_t)sc.match, sc_mis = (int8_t)-sc.mismatch;
    const bool bLeft = !( flag & MA_KSW_RIGHT );
    const int qe_row0 = q + e; // sic: the scalar qe of H[0] = v[0] - qe is initialised BEFORE the swap (:338, :247)
    if( q2 + e2 < q + e ) // kswcpp_core.h:367-375
        std::swap( q, q2 ), std::swap( e, e2 );
    const int qe = q + e, qe2 = q2 + e2;
    if( w < 0 )
        w = std::max( tlen, qlen );
    const int64_t T16 = ( ( tlen + 15 ) / 16 ) * 16, Q16 = ( ( qlen + 15 ) / 16 ) * 16;
    int64_t n_col = std::min( qlen, tlen );
    n_col = ( ( std::min<int64_t>( n_col, w + 1 ) + 15 ) / 16 + 1 ) * 16; // in cells
    // min/max over the 5x5 matrix {match, -mismatch, 0}  (kswcpp_core.h:406-412, kswcpp.h:85-95)
    int max_sc = std::max<int>( sc_mch, 0 ), min_sc = std::min<int>( sc_mis, 0 );
    (void)max_sc;
    if( -min_sc > 2 * ( q + e ) )
        return;
    int64_t long_thres = e != e2 ? ( q2 - q ) / ( e - e2 ) - 1 : 0; // :414-417
    if( q2 + e2 + long_thres * e2 > q + e + long_thres * e )
        ++long_thres;
    const int64_t long_diff = long_thres * ( e - e2 ) - ( q2 - q ) - e2;

    std::vector<int8_t> u( T16 + 32, w8( -q - e ) ), v( T16 + 32, w8( -q - e ) ), x( T16 + 32, w8( -q - e ) ),
        y( T16 + 32, w8( -q - e ) ), x2( T16 + 32, w8( -q2 - e2 ) ), y2( T16 + 32, w8( -q2 - e2 ) ), s( T16 + 32, 0 );
    std::vector<TS> H( T16 + 32, NEG_INF );
    // sf = target copy, zero padded to T16 and followed in memory by qr = reversed query, zero padded (:434-444, :508-514)
    std::vector<uint8_t> mem( T16 + Q16 + 64, 0 );
    uint8_t* sf = mem.data( );
    uint8_t* qr = sf + T16;
    for( int t = 0; t < qlen; t++ )
        qr[ t ] = query[ qlen - 1 - t ];
    memcpy( sf, target, tlen );
    std::vector<uint8_t> p( (size_t)( qlen + tlen - 1 ) * n_col + 16 );
    std::vector<int> off( qlen + tlen, 0 ), off_end( qlen + tlen, 0 );

    int64_t last_st = -1, last_en = -1, cells = 0;
    for( int64_t r = 0; r < qlen + tlen - 1; ++r )
    {
        int64_t st = 0, en = tlen - 1; // :541-553
        if( st < r - qlen + 1 )
            st = r - qlen + 1;
        if( en > r )
            en = r;
        if( st < ( ( r - w + 1 ) >> 1 ) )
            st = ( r - w + 1 ) >> 1;
        if( en > ( ( r + w ) >> 1 ) )
            en = ( r + w ) >> 1;
        if( st > en )
        {
            ez->zdropped = 1;
            break;
        }
        const int64_t st0 = st, en0 = en;
        cells += en0 - st0 + 1;
        st = ( st / 16 ) * 16;
        en = ( en + 16 ) / 16 * 16 - 1;
        int8_t x1, x21, v1; // :562-585
        const int8_t first_col = w8( r == 0 ? -q - e : r < long_thres ? -e : r == long_thres ? long_diff : -e2 );
        if( st > 0 )
        {
            if( st - 1 >= last_st && st - 1 <= last_en )
                x1 = x[ st - 1 ], x21 = x2[ st - 1 ], v1 = v[ st - 1 ];
            else
                x1 = w8( -q - e ), x21 = w8( -q2 - e2 ), v1 = w8( -q - e );
        }
        else
            x1 = w8( -q - e ), x21 = w8( -q2 - e2 ), v1 = first_col;
        if( en >= r )
        {
            y[ r ] = w8( -q - e );
            y2[ r ] = w8( -q2 - e2 );
            u[ r ] = first_col;
        }
        // score profile, 16 cells at a time starting at the UNaligned st0 (:591-616); N scores -e2
        {
            const uint8_t* qrr = qr + ( qlen - 1 - r ); // may point before qr for r >= qlen: never dereferenced there
            for( int64_t t = st0; t <= en0; t += 16 )
                for( int l = 0; l < 16; l++ )
                {
                    int64_t tt = t + l;
                    if( tt >= T16 )
                        break; // the reference spills into sf[0..14] here; those bytes are never read again
                    uint8_t a = sf[ tt ], b = qrr[ tt ];
                    s[ tt ] = ( a == 4 || b == 4 ) ? w8( -e2 ) : ( a == b ? sc_mch : sc_mis );
                }
        }
        off[ r ] = (int)st;
        off_end[ r ] = (int)en;
        uint8_t* pr = p.data( ) + r * n_col - st;
        for( int64_t t = st; t <= en; ++t ) // :653-766, element-wise view of the vector code
        {
            int8_t z = s[ t ];
            const int8_t xt1 = x1, vt1 = v1, x2t1 = x21;
            x1 = x[ t ], v1 = v[ t ], x21 = x2[ t ]; // old values move on to t+1
            const int8_t ut = u[ t ];
            int8_t a = w8( xt1 + vt1 ), b = w8( y[ t ] + ut ), a2 = w8( x2t1 + vt1 ), b2 = w8( y2[ t ] + ut );
            uint8_t d;
            if( bLeft )
            {
                d = a > z ? 1 : 0;
                z = std::max( z, a );
                d = b > z ? 2 : d;
                z = std::max( z, b );
                d = a2 > z ? 3 : d;
                z = std::max( z, a2 );
                d = b2 > z ? 4 : d;
                z = std::max( z, b2 );
            }
            else
            { // right-aligned gaps: ties go to the gap; state 4 is never recorded (:693-699)
                d = z > a ? 0 : 1;
                z = std::max( z, a );
                d = z > b ? d : 2;
                z = std::max( z, b );
                d = z > a2 ? d : 3;
                z = std::max( z, a2 );
                z = std::max( z, b2 );
            }
            z = std::min( z, sc_mch );
            u[ t ] = w8( z - vt1 );
            v[ t ] = w8( z - ut );
            int8_t tmp = w8( z - q );
            a = w8( a - tmp ), b = w8( b - tmp );
            tmp = w8( z - q2 );
            a2 = w8( a2 - tmp ), b2 = w8( b2 - tmp );
            if( bLeft )
            {
                x[ t ] = w8( ( a > 0 ? a : 0 ) - qe ), d |= a > 0 ? 0x08 : 0;
                y[ t ] = w8( ( b > 0 ? b : 0 ) - qe ), d |= b > 0 ? 0x10 : 0;
                x2[ t ] = w8( ( a2 > 0 ? a2 : 0 ) - qe2 ), d |= a2 > 0 ? 0x20 : 0;
                y2[ t ] = w8( ( b2 > 0 ? b2 : 0 ) - qe2 ), d |= b2 > 0 ? 0x40 : 0;
            }
            else
            {
                x[ t ] = w8( ( 0 > a ? 0 : a ) - qe ), d |= 0 > a ? 0 : 0x08;
                y[ t ] = w8( ( 0 > b ? 0 : b ) - qe ), d |= 0 > b ? 0 : 0x10;
                x2[ t ] = w8( ( 0 > a2 ? 0 : a2 ) - qe2 ), d |= 0 > a2 ? 0 : 0x20;
                y2[ t ] = w8( ( 0 > b2 ? 0 : b2 ) - qe2 ), d |= 0 > b2 ? 0 : 0x40;
            }
            pr[ t ] = d;
        }
        // calcMaxScore, exact branch (:178-264). Lane-blocked arg-max: SIZE lanes, block base recorded, not t+lane.
        TS max_H, max_t;
        if( r > 0 )
        {
            const int64_t en1 = st0 + ( ( en0 - st0 ) / SIZE ) * SIZE;
            max_H = H[ en0 ] = (TS)( en0 > 0 ? H[ en0 - 1 ] + u[ en0 ] : H[ en0 ] + v[ en0 ] );
            max_t = (TS)en0;
            TS lane_H[ SIZE ], lane_t[ SIZE ];
            for( int l = 0; l < SIZE; l++ )
                lane_H[ l ] = max_H, lane_t[ l ] = max_t;
            TS t;
            for( t = (TS)st0; t < (TS)en1; t += SIZE )
                for( int l = 0; l < SIZE; l++ )
                {
                    H[ t + l ] = (TS)( H[ t + l ] + v[ t + l ] );
                    if( H[ t + l ] > lane_H[ l ] )
                        lane_H[ l ] = H[ t + l ], lane_t[ l ] = t;
                }
            max_H = *std::max_element( lane_H, lane_H + SIZE );
            max_t = *std::max_element( lane_t, lane_t + SIZE );
            for( ; t < (TS)en0; ++t )
            {
                H[ t ] = (TS)( H[ t ] + v[ t ] );
                if( H[ t ] > max_H )
                    max_H = H[ t ], max_t = t;
            }
        }
        else
        {
            H[ 0 ] = (TS)( v[ 0 ] - qe_row0 );
            max_H = H[ 0 ];
            max_t = 0;
        }
        int64_t escB = std::numeric_limits<int64_t>::min( );
        if( g_pEsc )
            for( int64_t t = st0; t <= en0; t++ )
                escB = std::max<int64_t>( escB, (int64_t)H[ t ] + (int64_t)sc.match * ( qlen - 1 - ( r - t ) ) );
        if( en0 == tlen - 1 && H[ en0 ] > ez->mte )
            ez->mte = H[ en0 ], ez->mte_q = (int)( r - en ); // sic: the 16-aligned en (:254-255, :774)
        if( r - st0 == qlen - 1 && H[ st0 ] > ez->mqe )
            ez->mqe = H[ st0 ], ez->mqe_t = (int)st0;
        { // ksw_apply_zdrop(ez, 1, max_H, r, max_t, zdrop, e2)  (:22-44) — e2 AFTER the q/q2 swap
            const int tt = max_t, rr = (int)r;
            const int32_t Hh = max_H;
            bool bStop = false;
            if( Hh > (int32_t)ez->max )
            {
                if( g_pEsc && g_pEsc->stop_row >= 0 )
                    g_pEsc->violated = 1;
                ez->max = Hh, ez->max_t = tt, ez->max_q = rr - tt;
            }
            else if( tt >= ez->max_t && rr - tt >= ez->max_q )
            {
                int tl = tt - ez->max_t, ql = ( rr - tt ) - ez->max_q, l;
                l = tl > ql ? tl - ql : ql - tl;
                if( zdrop >= 0 && ( int32_t )( ez->max - Hh ) > zdrop + l * e2 )
                {
                    ez->zdropped = 1;
                    bStop = true;
                }
            }
            if( bStop )
                break;
        }
        if( r == qlen + tlen - 2 && en0 == tlen - 1 )
            ez->score = H[ tlen - 1 ];
        if( g_pEsc )
        {
            g_pEsc->rows = r + 1;
            const int64_t j = r + 1;
            const int64_t mingap = std::min<int64_t>( (int64_t)q + (int64_t)e * j, (int64_t)q2 + (int64_t)e2 * j );
            const int64_t T = (int64_t)sc.match * qlen - mingap;
            if( g_pEsc->stop_row < 0 && r >= qlen && escPrevB != std::numeric_limits<int64_t>::max( ) &&
                std::max( { escB, escPrevB, T } ) <= (int64_t)ez->max )
                g_pEsc->stop_row = r;
            escPrevB = escB;
        }
        last_st = st, last_en = en;
    }
    if( pCells )
        *pCells = cells;
    const bool rev_cigar = flag & MA_KSW_REV_CIGAR; // :796-835
    if( !ez->zdropped && !( flag & MA_KSW_EXTZ_ONLY ) )
        backtrack( rev_cigar, p, off, off_end, n_col, tlen - 1, qlen - 1, cig );
    else if( !ez->zdropped && ( flag & MA_KSW_EXTZ_ONLY ) && ez->mqe > (int)ez->max )
    {
        ez->reach_end = 1;
        backtrack( rev_cigar, p, off, off_end, n_col, ez->mqe_t, qlen - 1, cig );
    }
    else if( ez->max_t >= 0 && ez->max_q >= 0 )
        backtrack( rev_cigar, p, off, off_end, n_col, ez->max_t, ez->max_q, cig );
}
} // namespace

// returns rows processed by the reference; *stop_row = row after which the product may stop (-1: never),
// *violated = 1 if a later row still raised ez.max (must never happen)
extern "C" int ma_oracle_ksw_earlystop_check( int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                              const ma_oracle_score_t* sc, int w, int zdrop, int flag,
                                              int64_t* stop_row, int64_t* violated, int64_t* rows );

extern "C" int ma_oracle_ksw( int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                              const ma_oracle_score_t* sc, int w, int zdrop, int flag, ma_oracle_ksw_t* ez,
                              uint32_t* cigar, int cigar_cap, int64_t* cells )
{
    // ksw_reset_extz (kswcpp_core.h:15-20)
    ez->max_q = ez->max_t = ez->mqe_t = ez->mte_q = -1;
    ez->max = 0;
    ez->score = ez->mqe = ez->mte = std::numeric_limits<int>::min( );
    ez->n_cigar = 0, ez->zdropped = 0, ez->reach_end = 0;
    if( cells )
        *cells = 0;
    if( qlen <= 0 || tlen <= 0 )
        return 0;
    Cigar cig;
    // kswcpp.h:101-115 with iOverallMinScr = min(-mismatch, -gap, -extend, -gap2, -extend2)
    const int64_t iSize = std::max( qlen, tlen );
    const int64_t iMin = std::min( { -sc->mismatch, -sc->gap, -sc->extend, -sc->gap2, -sc->extend2 } );
    const bool bRisk16 = iSize * iMin < -32768 || iSize * sc->match > 32767;
    if( !bRisk16 )
        core<int16_t, 8>( qlen, query, tlen, target, *sc, w, zdrop, flag, ez, cig, cells );
    else
        core<int32_t, 4>( qlen, query, tlen, target, *sc, w, zdrop, flag, ez, cig, cells );
    ez->n_cigar = (int)cig.v.size( );
    if( (int)cig.v.size( ) > cigar_cap )
        return -1;
    if( !cig.v.empty( ) )
        memcpy( cigar, cig.v.data( ), cig.v.size( ) * 4 );
    return 0;
}

extern "C" int ma_oracle_ksw_earlystop_check( int qlen, const uint8_t* query, int tlen, const uint8_t* target,
                                              const ma_oracle_score_t* sc, int w, int zdrop, int flag,
                                              int64_t* stop_row, int64_t* violated, int64_t* rows )
{
    EarlyStopCheck esc;
    g_pEsc = &esc;
    ma_oracle_ksw_t ez;
    std::vector<uint32_t> cig( (size_t)qlen + tlen + 8 );
    int64_t cells;
    int rc = ma_oracle_ksw( qlen, query, tlen, target, sc, w, zdrop, flag, &ez, cig.data( ), (int)cig.size( ), &cells );
    g_pEsc = nullptr;
    *stop_row = esc.stop_row, *violated = esc.violated, *rows = esc.rows;
    return rc;
}
