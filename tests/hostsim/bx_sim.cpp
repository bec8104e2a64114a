// Test infrastructure: runs the packed banded exact DP kernel source (ma_b200/csrc/ksw_bx.cuh) on the lock-step warp
// emulator and compares EVERY kswcpp_extz_t field and the CIGAR with the oracle's kswcpp computation.
//   bx_sim <n_problems> <seed> [maxlen]     exit code 0 = all identical
#include "warp_emu.h"
#define MA_WARP_EMU 1
#include "../../ma_b200/csrc/ksw_bn.cuh"
#include "../../oracle/oracle.h"
#include <algorithm>
#include <random>
#include <string>
#include <vector>

using namespace ma;

static KswScore makeScore( int match, int mismatch, int gap, int extend, int gap2, int extend2 )
{ // ma_b200.cu make_score
    KswScore s;
    s.match = match, s.mismatch = -mismatch;
    int q = gap, e = extend, q2 = gap2, e2 = extend2;
    s.qe_row0 = q + e;
    if( q2 + e2 < q + e )
        std::swap( q, q2 ), std::swap( e, e2 );
    s.q = q, s.e = e, s.q2 = q2, s.e2 = e2;
    long long lt = e != e2 ? ( q2 - q ) / ( e - e2 ) - 1 : 0;
    if( q2 + e2 + lt * e2 > q + e + lt * e )
        ++lt;
    s.long_thres = (int)lt;
    s.long_diff = (int)( lt * ( e - e2 ) - ( q2 - q ) - e2 );
    s.min16 = std::min( { -mismatch, -gap, -extend, -gap2, -extend2 } );
    int min_sc = std::min( -mismatch, 0 );
    s.early_return = ( -min_sc > 2 * ( q + e ) ) ? 1 : 0;
    return s;
}

static bool g_bNarrow = false; // BX_NARROW=1: problems of at most three chunks run ksw_bn_rows

struct Result
{
    bool ok;
    KswOut ez;
    std::vector<unsigned> cigar;
};

template <int W, bool LEFT>
static Result runBx( const KswScore& P, const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int w, int zdrop,
                     int flag, bool bEarly )
{
    Result R;
    std::vector<uint8_t> slab( q );
    slab.insert( slab.end( ), t.begin( ), t.end( ) );
    SeqAccess sa;
    sa.qbase = slab.data( ), sa.qoff = 0, sa.qstep = 1, sa.tslab = slab.data( ), sa.toff = (long long)q.size( ), sa.tstep = 1;
    sa.pac = nullptr, sa.fwd_len = 0;
    const int qlen = (int)q.size( ), tlen = (int)t.size( );
    const int nc = ksw_ncol16( qlen, tlen, w );
    std::vector<unsigned char> tb( (size_t)( qlen + tlen ) * nc + 64, 0xEE );
    std::vector<unsigned> cs( (size_t)qlen + tlen + 8 );
    static KswBxSmem<W> sm;
    static KswBnSmem smn;
    memset( &sm, 0xA5, sizeof( sm ) );
    memset( &smn, 0xA5, sizeof( smn ) );
    const int nch = ( nc + 63 ) / 64; // register-resident narrow-band kernel (ksw_bn.cuh) where it applies
    KswOut outs[ 32 ];
    bool ok[ 32 ];
    int ncig = 0;
    warpemu::run( [ & ]( ) {
        KswOut ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
        ez.max = 0;
        ez.score = ez.mqe = ez.mte = (int)0x80000000;
        ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.status = 0, ez.cells = 0, ez.cigar_off = 0;
        const BxK K = ksw_bx_make_k( P, LEFT );
        if( g_bNarrow && nch == 1 )
            ksw_bn_rows<1, LEFT>( K, P, sa, qlen, tlen, w, zdrop, bEarly, smn, tb.data( ), ez );
        else if( g_bNarrow && nch == 2 )
            ksw_bn_rows<2, LEFT>( K, P, sa, qlen, tlen, w, zdrop, bEarly, smn, tb.data( ), ez );
        else if( g_bNarrow && nch == 3 )
            ksw_bn_rows<3, LEFT>( K, P, sa, qlen, tlen, w, zdrop, bEarly, smn, tb.data( ), ez );
        else
            ksw_bx_rows<W, LEFT>( K, P, sa, qlen, tlen, w, zdrop, bEarly, sm, tb.data( ), ez );
        const bool b = true;
        const int lane = warpemu::lane( );
        ok[ lane ] = b;
        int i0 = 0, j0 = 0;
        const bool bBt = ksw_bt_start( ez, qlen, tlen, flag, i0, j0 );
        outs[ lane ] = ez;
        if( lane == 0 && bBt )
            ncig = ksw_backtrack( tb.data( ), nc, qlen, tlen, w, i0, j0, cs.data( ), (int)cs.size( ) );
    } );
    for( int l = 1; l < 32; l++ )
        if( ok[ l ] != ok[ 0 ] || memcmp( &outs[ l ], &outs[ 0 ], sizeof( KswOut ) ) != 0 )
        {
            fprintf( stderr, "lanes disagree\n" );
            exit( 3 );
        }
    R.ok = ok[ 0 ];
    R.ez = outs[ 0 ];
    const bool rev = flag & MA_KSW_REV_CIGAR;
    for( int k = 0; k < ncig; k++ )
        R.cigar.push_back( rev ? cs[ k ] : cs[ ncig - 1 - k ] );
    return R;
}

template <bool LEFT>
static Result runW( const KswScore& P, const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int w, int zdrop, int flag,
                    bool bEarly )
{
    const int need = ksw_ncol16( (int)q.size( ), (int)t.size( ), w ) + 48;
    if( need <= 128 )
        return runBx<128, LEFT>( P, q, t, w, zdrop, flag, bEarly );
    if( need <= 256 )
        return runBx<256, LEFT>( P, q, t, w, zdrop, flag, bEarly );
    if( need <= 512 )
        return runBx<512, LEFT>( P, q, t, w, zdrop, flag, bEarly );
    if( need <= 640 ) // the band-512 class: a window that is not a power of two
        return runBx<640, LEFT>( P, q, t, w, zdrop, flag, bEarly );
    return runBx<1024, LEFT>( P, q, t, w, zdrop, flag, bEarly );
}

static bool compare( const Result& R, const int sc[ 6 ], const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int w,
                     int zdrop, int flag, bool bEarly, long long* cigOps, const char* what )
{
    const int ql = (int)q.size( ), tl = (int)t.size( );
    ma_oracle_score_t os{ sc[ 0 ], sc[ 1 ], sc[ 2 ], sc[ 3 ], sc[ 4 ], sc[ 5 ] };
    ma_oracle_ksw_t oz;
    std::vector<uint32_t> oc( (size_t)ql + tl + 8 );
    int64_t cells = 0;
    ma_oracle_ksw( ql, q.data( ), tl, t.data( ), &os, w, zdrop, flag, &oz, oc.data( ), (int)oc.size( ), &cells );
    bool same = R.ez.max == oz.max && R.ez.max_q == oz.max_q && R.ez.max_t == oz.max_t;
    bool cigComparable = true;
    if( bEarly ) // early-stop callers consume max / max_q / max_t and the CIGAR from that position
        cigComparable = oz.zdropped || !( oz.mqe > oz.max );
    else
        same = same && R.ez.zdropped == oz.zdropped && R.ez.mqe == oz.mqe && R.ez.mqe_t == oz.mqe_t && R.ez.mte == oz.mte &&
               R.ez.mte_q == oz.mte_q && R.ez.score == oz.score && R.ez.reach_end == oz.reach_end && R.ez.cells == cells;
    if( same && cigComparable )
    {
        same = (int)R.cigar.size( ) == oz.n_cigar;
        for( int k = 0; same && k < oz.n_cigar; k++ )
            same = R.cigar[ k ] == oc[ k ];
        *cigOps += oz.n_cigar;
    }
    if( !same )
        fprintf( stderr,
                 "MISMATCH %s ql=%d tl=%d w=%d zd=%d flag=%x early=%d:\n got max %d q %d t %d zd %d mqe %d "
                 "mqe_t %d mte %d mte_q %d score %d re %d cells %lld ncig %zu\n ref max %d q %d t %d zd %d mqe %d mqe_t %d "
                 "mte %d mte_q %d score %d re %d cells %lld ncig %d\n",
                 what, ql, tl, w, zdrop, flag, (int)bEarly, R.ez.max, R.ez.max_q, R.ez.max_t, R.ez.zdropped, R.ez.mqe,
                 R.ez.mqe_t, R.ez.mte, R.ez.mte_q, R.ez.score, R.ez.reach_end, R.ez.cells, R.cigar.size( ), oz.max, oz.max_q,
                 oz.max_t, oz.zdropped, oz.mqe, oz.mqe_t, oz.mte, oz.mte_q, oz.score, oz.reach_end, (long long)cells,
                 oz.n_cigar );
    return same;
}

// bx_sim file <path> m x g e g2 e2: lines "w zdrop flag query target" (digits)
static int runFile( int argc, char** argv )
{
    FILE* f = fopen( argv[ 2 ], "r" );
    if( !f || argc < 9 )
        return 2;
    int sc[ 6 ];
    for( int k = 0; k < 6; k++ )
        sc[ k ] = atoi( argv[ 3 + k ] );
    const KswScore P = makeScore( sc[ 0 ], sc[ 1 ], sc[ 2 ], sc[ 3 ], sc[ 4 ], sc[ 5 ] );
    if( !ksw_bx_params_ok( P ) || P.early_return )
    {
        printf( "bx_sim file: scores outside the packed mode\n" );
        return 0;
    }
    static char qb[ 1 << 16 ], tbuf[ 1 << 16 ];
    int w, zd, fl, n = 0, bad = 0, bail = 0;
    long long cigOps = 0;
    while( fscanf( f, "%d %d %d %65535s %65535s", &w, &zd, &fl, qb, tbuf ) == 5 )
    {
        std::vector<uint8_t> q, t;
        for( char* c = qb; *c; c++ )
            q.push_back( (uint8_t)( *c - '0' ) );
        for( char* c = tbuf; *c; c++ )
            t.push_back( (uint8_t)( *c - '0' ) );
        if( w < 0 )
            w = (int)std::max( q.size( ), t.size( ) );
        if( ksw_ncol16( (int)q.size( ), (int)t.size( ), w ) + 48 > 1024 )
            continue;
        const bool left = !( fl & MA_KSW_RIGHT );
        const Result R = left ? runW<true>( P, q, t, w, zd, fl, false ) : runW<false>( P, q, t, w, zd, fl, false );
        n++;
        if( !R.ok )
        {
            bail++;
            continue;
        }
        char what[ 32 ];
        snprintf( what, sizeof what, "line %d", n - 1 );
        bad += !compare( R, sc, q, t, w, zd, fl, false, &cigOps, what );
    }
    fclose( f );
    printf( "bx_sim file: %d problems, %d handed over, %lld cigar ops compared, %d MISMATCHES\n", n, bail,
            cigOps, bad );
    return bad ? 1 : 0;
}

int main( int argc, char** argv )
{
    g_bNarrow = getenv( "BX_NARROW" ) != nullptr;
    if( argc > 2 && std::string( argv[ 1 ] ) == "file" )
        return runFile( argc, argv );
    const int n = argc > 1 ? atoi( argv[ 1 ] ) : 200;
    const unsigned seed = argc > 2 ? (unsigned)atoi( argv[ 2 ] ) : 1;
    const int maxlen = argc > 3 ? atoi( argv[ 3 ] ) : 260;
    std::mt19937_64 rng( seed );
    auto rnd = [ & ]( int lo, int hi ) { return lo + (int)( rng( ) % (unsigned long long)( hi - lo + 1 ) ); };
    int nRun = 0, nBad = 0, nZdrop = 0, nEarly = 0, nBail = 0;
    long long cigOps = 0;
    for( int it = 0; it < n; it++ )
    {
        int sc[ 6 ] = { 2, 4, 4, 2, 24, 1 };
        const int sk = it % 8;
        if( sk == 5 )
            sc[ 0 ] = 1, sc[ 1 ] = 4, sc[ 2 ] = 6, sc[ 3 ] = 1, sc[ 4 ] = 6, sc[ 5 ] = 1; // bwa-like, e == e2
        if( sk == 6 )
            sc[ 0 ] = 2, sc[ 1 ] = 4, sc[ 2 ] = 4, sc[ 3 ] = 2, sc[ 4 ] = 3, sc[ 5 ] = 1; // swapped pieces, negative long-gap threshold
        if( sk == 7 )
            sc[ 0 ] = 2, sc[ 1 ] = 3, sc[ 2 ] = 2, sc[ 3 ] = 3, sc[ 4 ] = 10, sc[ 5 ] = 2;
        if( sk == 4 ) // beyond the range in which the reference's int8 values cannot wrap
            sc[ 0 ] = 5, sc[ 1 ] = 9, sc[ 2 ] = 12, sc[ 3 ] = 3, sc[ 4 ] = 40, sc[ 5 ] = 2;
        if( sk == 3 && it % 16 == 3 )
            sc[ 0 ] = 11, sc[ 1 ] = 25, sc[ 2 ] = 33, sc[ 3 ] = 7, sc[ 4 ] = 90, sc[ 5 ] = 3;
        const KswScore P = makeScore( sc[ 0 ], sc[ 1 ], sc[ 2 ], sc[ 3 ], sc[ 4 ], sc[ 5 ] );
        if( !ksw_bx_params_ok( P ) || P.early_return )
            continue;
        const int kind = rnd( 0, 6 );
        const int ql = kind == 6 ? rnd( 1, 12 ) : rnd( 1, maxlen );
        int tl = kind == 6 ? rnd( 1, 40 ) : std::max( 1, ql + rnd( -ql / 2, ql / 2 + 40 ) );
        std::vector<uint8_t> t( tl ), q;
        const int alpha = kind == 3 ? 2 : 4;
        for( auto& c : t )
            c = (uint8_t)rnd( 0, alpha - 1 );
        if( kind == 1 )
        { // tandem repeat
            const int per = rnd( 1, 12 );
            for( int i = per; i < tl; i++ )
                t[ i ] = t[ i - per ];
        }
        // the query: the target with substitutions, insertions and deletions
        const int rate[ 4 ] = { 0, 2, 10, 30 };
        const int mr = rate[ rnd( 0, 3 ) ], ir = rate[ rnd( 0, 2 ) ];
        for( int i = 0; (int)q.size( ) < ql; i++ )
        {
            if( rnd( 0, 99 ) < ir )
            {
                if( rnd( 0, 1 ) )
                    for( int k = rnd( 1, 12 ); k > 0 && (int)q.size( ) < ql; k-- )
                        q.push_back( (uint8_t)rnd( 0, 3 ) );
                else
                    i += rnd( 1, 12 );
            }
            uint8_t c = kind == 4 ? (uint8_t)rnd( 0, 3 ) : t[ i % tl ];
            if( rnd( 0, 99 ) < mr )
                c = (uint8_t)( ( c + 1 ) & 3 );
            if( (int)q.size( ) < ql )
                q.push_back( c );
        }
        if( rnd( 0, 9 ) == 0 )
            q[ rnd( 0, ql - 1 ) ] = 4;
        if( rnd( 0, 9 ) == 0 )
            t[ rnd( 0, tl - 1 ) ] = 4;
        const int wsel = rnd( 0, 9 );
        int w = wsel == 0 ? -1 : wsel == 1 ? rnd( 0, 3 ) : wsel < 5 ? rnd( 4, 40 ) : wsel < 8 ? rnd( 41, 200 ) : rnd( 201, 900 );
        if( w < 0 )
            w = std::max( ql, tl );
        if( ksw_ncol16( ql, tl, w ) + 48 > 1024 )
            w = 100;
        const int zdrop = ( it % 3 == 0 ) ? -1 : ( it % 3 == 1 ) ? rnd( 5, 60 ) : 200;
        const int fsel = rnd( 0, 3 );
        const int flag = fsel == 0 ? 0 : fsel == 1 ? MA_KSW_RIGHT : fsel == 2 ? MA_KSW_EXTZ_ONLY
                                                                             : ( MA_KSW_EXTZ_ONLY | MA_KSW_RIGHT | MA_KSW_REV_CIGAR );
        const bool left = !( flag & MA_KSW_RIGHT );
        const bool bEarly = ( flag & MA_KSW_EXTZ_ONLY ) && rnd( 0, 2 ) == 0;
        const Result R = left ? runW<true>( P, q, t, w, zdrop, flag, bEarly ) : runW<false>( P, q, t, w, zdrop, flag, bEarly );
        if( !R.ok )
        {
            nBail++;
            continue;
        }
        nRun++;
        nZdrop += R.ez.zdropped;
        nEarly += bEarly;
        char what[ 64 ];
        snprintf( what, sizeof what, "it=%d kind=%d sk=%d", it, kind, sk );
        nBad += !compare( R, sc, q, t, w, zdrop, flag, bEarly, &cigOps, what );
    }
    printf( "bx_sim: %d problems, %d run (%d early-stop), %d handed over, %d z-dropped, %lld cigar ops "
            "compared, %d MISMATCHES\n",
            n, nRun, nEarly, nBail, nZdrop, cigOps, nBad );
    return nBad ? 1 : 0;
}
