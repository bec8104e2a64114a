// Test infrastructure: runs the query-stationary DP kernel source (ma_b200/csrc/ksw_qs.cuh) on the lock-step warp
// emulator and compares max / max_q / max_t / CIGAR with the oracle's full kswcpp computation.
//   qs_sim <n_problems> <seed>      exit code 0 = all identical
#include "warp_emu.h"
#define MA_WARP_EMU 1
#include "../../ma_b200/csrc/ksw_qs.cuh"
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

struct Result
{
    bool ok;
    KswOut ez;
    std::vector<unsigned> cigar;
};

template <int NB, bool LEFT>
static Result runQs( const KswScore& P, const std::vector<uint8_t>& q, const std::vector<uint8_t>& t, int w, int zdrop,
                     int flag )
{
    Result R;
    std::vector<uint8_t> slab( q );
    slab.insert( slab.end( ), t.begin( ), t.end( ) );
    SeqAccess sa;
    sa.qbase = slab.data( ), sa.qoff = 0, sa.qstep = 1, sa.tslab = slab.data( ), sa.toff = (long long)q.size( ), sa.tstep = 1;
    sa.pac = nullptr, sa.fwd_len = 0;
    const int qlen = (int)q.size( ), tlen = (int)t.size( );
    const int rows = std::min( w + 2, qlen + tlen ) + 1;
    std::vector<unsigned char> tb( (size_t)rows * 64 * NB + 64, 0xEE );
    std::vector<unsigned> cs( (size_t)qlen + tlen + 8 );
    static KswQsSmem<NB> sm;
    bool ok[ 32 ];
    KswOut outs[ 32 ];
    int ncig = 0;
    warpemu::run( [ & ]( ) {
        KswOut ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
        ez.max = 0;
        ez.score = ez.mqe = ez.mte = (int)0x80000000;
        ez.n_cigar = 0, ez.zdropped = 0, ez.reach_end = 0, ez.status = 0, ez.cells = 0;
        const QsK K = ksw_qs_make_k( P, LEFT );
        const bool b = ksw_qs_rows<NB, LEFT>( K, P, sa, qlen, tlen, w, zdrop, sm, tb.data( ), ez );
        const int lane = warpemu::lane( );
        ok[ lane ] = b, outs[ lane ] = ez;
        if( b && lane == 0 && ez.max_t >= 0 && ez.max_q >= 0 )
            ncig = ksw_qs_backtrack( tb.data( ), 64 * NB, LEFT, qlen, tlen, w, ez.max_t, ez.max_q, cs.data( ), (int)cs.size( ) );
    } );
    for( int l = 1; l < 32; l++ )
        if( ok[ l ] != ok[ 0 ] || outs[ l ].max != outs[ 0 ].max || outs[ l ].max_t != outs[ 0 ].max_t ||
            outs[ l ].max_q != outs[ 0 ].max_q )
        {
            fprintf( stderr, "lanes disagree\n" );
            exit( 3 );
        }
    R.ok = ok[ 0 ], R.ez = outs[ 0 ];
    if( R.ok )
    {
        const bool rev = flag & MA_KSW_REV_CIGAR;
        for( int k = 0; k < ncig; k++ )
            R.cigar.push_back( rev ? cs[ k ] : cs[ ncig - 1 - k ] );
    }
    return R;
}

static int runFile( const char* path )
{ // lines: w zdrop flag query target (digits), default scores
    FILE* f = fopen( path, "r" );
    if( !f )
        return 2;
    const KswScore P = makeScore( 2, 4, 4, 2, 24, 1 );
    static char qb[ 1 << 16 ], tbuf[ 1 << 16 ];
    int w, zd, fl, n = 0, bad = 0, bail = 0;
    while( fscanf( f, "%d %d %d %65535s %65535s", &w, &zd, &fl, qb, tbuf ) == 5 )
    {
        std::vector<uint8_t> q, t;
        for( char* c = qb; *c; c++ )
            q.push_back( (uint8_t)( *c - '0' ) );
        for( char* c = tbuf; *c; c++ )
            t.push_back( (uint8_t)( *c - '0' ) );
        const int ql = (int)q.size( ), tl = (int)t.size( );
        const int nb = ksw_qs_class( P, ql, tl, w, MA_TASK_EARLYSTOP );
        const bool left = !( fl & MA_KSW_RIGHT );
        n++;
        if( nb == 0 )
            continue;
        Result R;
        if( left )
            R = nb == 1 ? runQs<1, true>( P, q, t, w, zd, fl ) : nb == 2 ? runQs<2, true>( P, q, t, w, zd, fl )
                                                                         : runQs<3, true>( P, q, t, w, zd, fl );
        else
            R = nb == 1 ? runQs<1, false>( P, q, t, w, zd, fl ) : nb == 2 ? runQs<2, false>( P, q, t, w, zd, fl )
                                                                          : runQs<3, false>( P, q, t, w, zd, fl );
        if( !R.ok )
        {
            bail++;
            continue;
        }
        ma_oracle_score_t os{ 2, 4, 4, 2, 24, 1 };
        ma_oracle_ksw_t oz;
        std::vector<uint32_t> oc( (size_t)ql + tl + 8 );
        int64_t cells = 0;
        ma_oracle_ksw( ql, q.data( ), tl, t.data( ), &os, w, zd, fl, &oz, oc.data( ), (int)oc.size( ), &cells );
        if( R.ez.max != oz.max || R.ez.max_q != oz.max_q || R.ez.max_t != oz.max_t )
        {
            bad++;
            fprintf( stderr, "MISMATCH problem %d: got max %d q %d t %d | ref max %d q %d t %d (qlen %d tlen %d)\n", n - 1,
                     R.ez.max, R.ez.max_q, R.ez.max_t, oz.max, oz.max_q, oz.max_t, ql, tl );
        }
    }
    fclose( f );
    printf( "qs_sim file: %d problems, %d handed over, %d MISMATCHES\n", n, bail, bad );
    return bad ? 1 : 0;
}

int main( int argc, char** argv )
{
    if( argc > 2 && std::string( argv[ 1 ] ) == "file" )
        return runFile( argv[ 2 ] );
    const int n = argc > 1 ? atoi( argv[ 1 ] ) : 200;
    const unsigned seed = argc > 2 ? (unsigned)atoi( argv[ 2 ] ) : 1;
    std::mt19937_64 rng( seed );
    auto rnd = [ & ]( int lo, int hi ) { return lo + (int)( rng( ) % (unsigned long long)( hi - lo + 1 ) ); };
    int nRun = 0, nBail = 0, nBad = 0, nZdrop = 0, nSkip = 0;
    long long cigOps = 0;
    for( int it = 0; it < n; it++ )
    {
        // scores: mostly the defaults, sometimes others that pass ksw_qs_params_ok (including the q / q2 swap)
        int sc[ 6 ] = { 2, 4, 4, 2, 24, 1 };
        const int sk = it % 8;
        if( sk == 5 )
            sc[ 0 ] = 1, sc[ 1 ] = 4, sc[ 2 ] = 6, sc[ 3 ] = 1, sc[ 4 ] = 6, sc[ 5 ] = 1; // bwa-like, e == e2
        if( sk == 6 )
            sc[ 0 ] = 3, sc[ 1 ] = 5, sc[ 2 ] = 20, sc[ 3 ] = 1, sc[ 4 ] = 3, sc[ 5 ] = 3; // swapped pieces
        if( sk == 7 )
            sc[ 0 ] = 2, sc[ 1 ] = 3, sc[ 2 ] = 2, sc[ 3 ] = 3, sc[ 4 ] = 10, sc[ 5 ] = 2;
        const KswScore P = makeScore( sc[ 0 ], sc[ 1 ], sc[ 2 ], sc[ 3 ], sc[ 4 ], sc[ 5 ] );
        const int kind = rnd( 0, 6 );
        const int ql = kind == 6 ? rnd( 1, 12 ) : rnd( 1, 192 );
        int tl = std::max( 2 * ql, 16 ) + rnd( 0, 600 );
        std::vector<uint8_t> t( tl ), q( ql );
        const int alpha = kind == 3 ? 2 : 4;
        for( auto& c : t )
            c = (uint8_t)rnd( 0, alpha - 1 );
        if( kind == 1 )
        { // tandem repeat
            const int per = rnd( 1, 12 );
            for( int i = per; i < tl; i++ )
                t[ i ] = t[ i - per ];
        }
        for( int i = 0; i < ql; i++ )
            q[ i ] = t[ i ];
        if( kind == 2 )
        { // a long deletion pays off
            const int cut = rnd( 0, ql - 1 ) + 1, far = rnd( 0, std::max( 1, tl - ql ) - 1 );
            for( int i = cut; i < ql; i++ )
                q[ i ] = t[ far + i - cut ];
        }
        if( kind == 4 )
            for( auto& c : q )
                c = (uint8_t)rnd( 0, 3 );
        if( kind == 5 )
        { // insertions in the query
            std::vector<uint8_t> q2;
            for( int i = 0; (int)q2.size( ) < ql; i++ )
            {
                if( rnd( 0, 9 ) == 0 )
                    for( int k = rnd( 1, 30 ); k > 0 && (int)q2.size( ) < ql; k-- )
                        q2.push_back( (uint8_t)rnd( 0, 3 ) );
                if( (int)q2.size( ) < ql )
                    q2.push_back( t[ i % tl ] );
            }
            q = q2;
        }
        const int mr = rnd( 0, 3 );
        const int rate[ 4 ] = { 0, 2, 10, 30 };
        for( auto& c : q )
            if( rnd( 0, 99 ) < rate[ mr ] )
                c = (uint8_t)( ( c + 1 ) & 3 );
        if( rnd( 0, 19 ) == 0 )
            q[ rnd( 0, ql - 1 ) ] = 4;
        const int w = ( it % 5 == 0 ) ? 2 * ql + 16 + rnd( 0, 40 ) : 512;
        const int zdrop = ( it % 7 == 0 ) ? rnd( 5, 60 ) : 200;
        const bool left = rnd( 0, 1 );
        const int flag = left ? MA_KSW_EXTZ_ONLY : ( MA_KSW_EXTZ_ONLY | MA_KSW_RIGHT | MA_KSW_REV_CIGAR );
        const int nb = ksw_qs_class( P, ql, tl, w, MA_TASK_EARLYSTOP );
        if( nb == 0 )
        {
            nSkip++;
            continue;
        }
        Result R;
        if( left )
            R = nb == 1 ? runQs<1, true>( P, q, t, w, zdrop, flag ) : nb == 2 ? runQs<2, true>( P, q, t, w, zdrop, flag )
                                                                             : runQs<3, true>( P, q, t, w, zdrop, flag );
        else
            R = nb == 1 ? runQs<1, false>( P, q, t, w, zdrop, flag ) : nb == 2 ? runQs<2, false>( P, q, t, w, zdrop, flag )
                                                                              : runQs<3, false>( P, q, t, w, zdrop, flag );
        if( !R.ok )
        {
            nBail++;
            continue;
        }
        nRun++;
        ma_oracle_score_t os{ sc[ 0 ], sc[ 1 ], sc[ 2 ], sc[ 3 ], sc[ 4 ], sc[ 5 ] };
        ma_oracle_ksw_t oz;
        std::vector<uint32_t> oc( (size_t)ql + tl + 8 );
        int64_t cells = 0;
        ma_oracle_ksw( ql, q.data( ), tl, t.data( ), &os, w, zdrop, flag, &oz, oc.data( ), (int)oc.size( ), &cells );
        nZdrop += R.ez.zdropped;
        // the reference backtracks from (max_t, max_q) unless it reached the end of the query without a z-drop
        // (reach_end): early-stop callers only consume max / max_q / max_t and the CIGAR from that position, which the
        // oracle reports when the extension z-dropped or mqe <= max
        bool same = R.ez.max == oz.max && R.ez.max_q == oz.max_q && R.ez.max_t == oz.max_t;
        const bool cigComparable = oz.zdropped || !( oz.mqe > oz.max );
        if( same && cigComparable )
        {
            same = (int)R.cigar.size( ) == oz.n_cigar;
            for( int k = 0; same && k < oz.n_cigar; k++ )
                same = R.cigar[ k ] == oc[ k ];
            cigOps += oz.n_cigar;
        }
        if( !same )
        {
            nBad++;
            if( nBad <= 10 )
                fprintf( stderr,
                         "MISMATCH it=%d kind=%d ql=%d tl=%d w=%d zd=%d left=%d sk=%d: got max %d q %d t %d ncig %zu | ref "
                         "max %d q %d t %d ncig %d zdropped %d mqe %d\n",
                         it, kind, ql, tl, w, zdrop, (int)left, sk, R.ez.max, R.ez.max_q, R.ez.max_t, R.cigar.size( ),
                         oz.max, oz.max_q, oz.max_t, oz.n_cigar, oz.zdropped, oz.mqe );
        }
    }
    printf( "qs_sim: %d problems, %d run, %d handed over, %d skipped, %d z-dropped, %lld cigar ops compared, %d MISMATCHES\n",
            n, nRun, nBail, nSkip, nZdrop, cigOps, nBad );
    return nBad ? 1 : 0;
}
