// CPU ORACLE — test infrastructure only (see oracle.h).
// Strip of Consideration and Harmonization restated with the same C++ runtime facilities the reference uses
// (std::sort, std::make_heap/pop_heap, glibc rand(), libm), so that their order-observable behaviour is inherited:
//   libs/ma/src/module/stripOfConsideration.cpp:12-161, libs/ma/inc/ma/container/soc.h:26-90, 196-284, 362-419
//   libs/ma/src/module/harmonization.cpp:14-173 (applyFilters), :182-249 (linesweep), :251-373 (harmonizeOne),
//   :374-555 (execute); libs/ma/inc/ma/module/harmonization.h:82-89 (deltaDistance)
//   libs/ma/src/sample_consensus/test_ransac.cpp:8-100, ransac.cpp:67-164, sac_model_line.cpp:49-131,
//   libs/ma/inc/ma/sample_consensus/lin_regres.h, test_ransac.h:20-74
#include "ma_oracle.h"
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>
#include <limits>

namespace oracle
{

// ------------------------------------------------------------------------------------------------ SoC
static void adjustScore( std::vector<Seed>& S, SoCOrder& rScore, size_t cutS, size_t cutE, size_t cntS, size_t cntE )
{ // soc.h:338-356
    if( (int64_t)cutE - (int64_t)cutS <= (int64_t)cntE - (int64_t)cntS )
        for( size_t i = cutS; i < cutE; i++ )
            rScore.sub( S[ i ] );
    else
    {
        SoCOrder n;
        for( size_t i = cntS; i < cntE; i++ )
            n.add( S[ i ] );
        rScore = n;
    }
}

static void push_back_no_overlap( SoCQueue& Q, SoCOrder cur, size_t itStrip, size_t itStripEnd, uint64_t uiMinScore )
{ // soc.h:362-404
    while( !Q.maxima.empty( ) && Q.maxima.back( ).end > itStrip )
    {
        SoC& b = Q.maxima.back( );
        if( b.order < cur )
        {
            adjustScore( Q.seeds, b.order, itStrip, b.end, b.begin, itStrip );
            b.end = itStrip;
            if( b.order.acc_len < uiMinScore || b.order.acc_len == 0 )
                Q.maxima.pop_back( );
        }
        else
        {
            adjustScore( Q.seeds, cur, itStrip, b.end, b.end, itStripEnd );
            itStrip = b.end;
            if( cur.acc_len < uiMinScore || cur.acc_len == 0 )
                return;
        }
    }
    Q.maxima.push_back( SoC{ cur, itStrip, itStripEnd } );
}

static bool heapOrder( const SoC& a, const SoC& b )
{
    return a.order < b.order;
}

SoCQueue strip_of_consideration( const Index& I, const Params& P, std::vector<Seed> seedsIn, int64_t qlen )
{
    SoCQueue Q;
    Q.seeds = std::move( seedsIn );
    auto& S = Q.seeds;
    if( S.empty( ) )
        return Q;
    double fMinLen = std::max( (double)P.harm_score_min_rel * qlen, (double)P.harm_score_min );
    if( (uint64_t)P.genome_size_disable >= (uint64_t)( 2 * I.fwd_len ) )
        fMinLen = 0;
    const uint64_t uiStripSize =
        P.soc_width != 0 ? (uint64_t)P.soc_width : (uint64_t)( ( P.match * qlen - P.gap ) / P.extend );
    std::sort( S.begin( ), S.end( ), []( const Seed& a, const Seed& b ) { return a.delta < b.delta; } );
    SoCOrder cur;
    size_t s = 0, e = 0;
    size_t cidS = (size_t)I.seqIdForPosition( S[ 0 ].r ), cidE = cidS;
    auto inContig = [ & ]( size_t id, int64_t pos ) {
        return I.contigs[ id ].start <= pos && pos < I.contigs[ id ].start + I.contigs[ id ].length;
    };
    while( e != S.size( ) && s != S.size( ) )
    {
        while( !inContig( cidS, S[ s ].r ) )
            cidS += 1; // rectangular: always counts upwards
        while( e != S.size( ) && (uint64_t)S[ s ].delta + uiStripSize >= (uint64_t)S[ e ].delta && cidS == cidE )
        {
            cur.add( S[ e ] );
            e++;
            if( e != S.size( ) )
                while( !inContig( cidE, S[ e ].r ) )
                    cidE += 1;
        }
        if( (double)cur.acc_len >= fMinLen )
            push_back_no_overlap( Q, cur, s, e, (uint64_t)fMinLen );
        cur.sub( S[ s ] );
        s++;
    }
    std::make_heap( Q.maxima.begin( ), Q.maxima.end( ), heapOrder );
    if( P.rectangular_soc )
    { // soc.h:196-231 — re-sorts the seeds and rewrites every window WITHOUT re-heapifying
        std::vector<std::pair<int64_t, int64_t>> vRef;
        for( auto& m : Q.maxima )
        {
            vRef.emplace_back( S[ m.begin ].r, S[ m.begin ].r );
            for( size_t i = m.begin; i != m.end; i++ )
            {
                vRef.back( ).first = std::min( vRef.back( ).first, S[ i ].r );
                vRef.back( ).second = std::max( vRef.back( ).second, S[ i ].r );
            }
        }
        std::sort( S.begin( ), S.end( ), []( const Seed& a, const Seed& b ) { return a.r < b.r; } );
        Q.maxima.clear( );
        for( auto& rp : vRef )
        {
            SoC m;
            m.begin = std::lower_bound( S.begin( ), S.end( ), rp.first,
                                        []( const Seed& s, int64_t p ) { return s.r < p; } ) -
                      S.begin( );
            size_t it = m.begin;
            while( it != S.size( ) && S[ it ].r <= rp.second )
                m.order.add( S[ it ] ), it++;
            m.end = it;
            Q.maxima.push_back( m );
        }
    }
    return Q;
}

std::vector<Seed> soc_pop( SoCQueue& Q, unsigned* pIndex ) // soc.h:240-284
{
    std::vector<Seed> ret;
    const SoC& f = Q.maxima.front( );
    *pIndex = Q.next_index++;
    for( size_t i = f.begin; i != Q.seeds.size( ) && i != f.end; i++ )
    {
        Q.seeds[ i ].soc_nt = f.order.acc_len;
        ret.push_back( Q.seeds[ i ] );
    }
    std::pop_heap( Q.maxima.begin( ), Q.maxima.end( ), heapOrder );
    Q.maxima.pop_back( );
    return ret;
}

// ------------------------------------------------------------------------------------------------ RANSAC
template <typename TP> static TP Median( std::vector<TP> arr ) // test_ransac.h:20-39
{
    std::sort( arr.begin( ), arr.end( ) );
    if( arr.size( ) == 0 )
        return 0;
    if( arr.size( ) == 1 )
        return arr[ 0 ];
    if( arr.size( ) % 2 == 0 )
        return ( arr[ arr.size( ) / 2 - 1 ] + arr[ arr.size( ) / 2 ] ) / 2;
    return arr[ arr.size( ) / 2 ];
}
static double medianAbsoluteDeviation( std::vector<double> arr ) // test_ransac.h:57-74
{
    double median = Median( arr );
    std::vector<double> dev;
    for( size_t i = 0; i < arr.size( ); i++ )
        if( arr[ i ] - median < 0 )
            dev.push_back( -( arr[ i ] - median ) );
        else
            dev.push_back( arr[ i ] - median );
    return Median( dev );
}

static std::pair<double, double> lin_regres( const std::vector<double>& vx, const std::vector<double>& vy )
{ // lin_regres.h:58-147
    const size_t n = vx.size( );
    std::vector<double> dx( n ), dy( n );
    double sum = 0;
    for( size_t i = 0; i < n; i++ )
        sum = sum + vx[ i ];
    const double mean_x = sum / (double)n;
    sum = 0;
    for( size_t i = 0; i < n; i++ )
        sum = sum + vy[ i ];
    const double mean_y = sum / (double)n;
    double sx = 0;
    for( size_t i = 0; i < n; i++ )
    {
        dx[ i ] = vx[ i ] - mean_x;
        sx = sx + dx[ i ] * dx[ i ];
    }
    for( size_t i = 0; i < n; i++ )
        dy[ i ] = vy[ i ] - mean_y;
    double sum_xy = 0;
    for( size_t i = 0; i < n; i++ )
        sum_xy = sum_xy + dx[ i ] * dy[ i ];
    const double slope = sum_xy / sx;
    const double intercept = mean_y - slope * mean_x;
    return { std::atan( slope ), -intercept / slope };
}

static std::pair<double, double> run_ransac( const std::vector<double>& X, const std::vector<double>& Y, double fMAD )
{ // test_ransac.cpp:8-100 + ransac.cpp:67-164 + sac_model_line.cpp:49-131
    const int N = (int)X.size( );
    int iterations = 0;
    int n_best = -INT_MAX;
    double k = 1.0;
    std::vector<int> best_inliers, inliers;
    bool bHaveModel = false;
    const double threshold = fMAD, probability = 0.99;
    const int max_iterations = 100;
    while( iterations < k )
    {
        // getSamples (sac_model_line.cpp:49-76)
        int s0, s1;
        {
            const double trand = N / ( RAND_MAX + 1.0 );
            int idx = (int)( rand( ) * trand );
            s0 = idx;
            int iter = 0;
            do
            {
                idx = (int)( rand( ) * trand );
                s1 = idx;
                iter++;
                if( iter > 1000 )
                    break;
                iterations++;
            } while( s1 == s0 );
            iterations--;
        }
        const double mc[ 6 ] = { X[ s0 ], Y[ s0 ], 0, X[ s1 ], Y[ s1 ], 0 };
        double dH = mc[ 0 ] - mc[ 3 ], dV = mc[ 1 ] - mc[ 4 ];
        if( dH <= 0 && dV <= 0 )
            dH *= -1, dV *= -1;
        double dAngle = -90;
        if( dH > 0 && dV > 0 )
            dAngle = atan( dV / dH ) * 180 / std::acos( -1 );
        if( dAngle >= 20 && dAngle <= 70 )
        {
            // selectWithinDistance (sac_model_line.cpp:86-131), 3-D cross product form with z = 0
            const double sqr_threshold = threshold * threshold;
            inliers.clear( );
            const double p3x = mc[ 3 ] - mc[ 0 ], p3y = mc[ 4 ] - mc[ 1 ], p3z = mc[ 5 ] - mc[ 2 ];
            for( int i = 0; i < N; i++ )
            {
                const double p4x = mc[ 3 ] - X[ i ], p4y = mc[ 4 ] - Y[ i ], p4z = mc[ 5 ] - 0.0;
                const double cx = p4y * p3z - p4z * p3y, cy = p4z * p3x - p4x * p3z, cz = p4x * p3y - p4y * p3x;
                const double sqr_distance = ( cx * cx + cy * cy + cz * cz ) / ( p3x * p3x + p3y * p3y + p3z * p3z );
                if( sqr_distance < sqr_threshold )
                    inliers.push_back( i );
            }
            const int n_inliers = (int)inliers.size( );
            if( n_inliers > n_best )
            {
                n_best = n_inliers;
                best_inliers = inliers;
                bHaveModel = true;
                const double w = (double)n_inliers / (double)N;
                double p_no_outliers = 1 - pow( w, 2.0 );
                p_no_outliers = std::max( std::numeric_limits<double>::epsilon( ), p_no_outliers );
                p_no_outliers = std::min( 1 - std::numeric_limits<double>::epsilon( ), p_no_outliers );
                k = log( 1 - probability ) / log( p_no_outliers );
            }
        }
        else
            continue;
        iterations += 1;
        if( iterations > max_iterations )
            break;
    }
    if( !bHaveModel )
        return { std::nan( "" ), std::nan( "" ) };
    std::vector<double> vX, vY;
    for( int i : best_inliers )
        vX.push_back( X[ i ] ), vY.push_back( Y[ i ] );
    return lin_regres( vX, vY );
}

// ------------------------------------------------------------------------------------------------ Harmonization
#define ORACLE_PI 3.14159265
static double deltaDistance( const Seed& s, double fAngle, int64_t uiRStart ) // harmonization.h:82-89
{
    double y = s.r + s.q / std::tan( ORACLE_PI / 2 - fAngle );
    double x = ( y - uiRStart ) * std::sin( fAngle );
    double x_1 = s.q / std::sin( ORACLE_PI / 2 - fAngle );
    return std::abs( x - x_1 );
}

struct Shadow
{
    size_t seed; // index into the seed vector
    uint64_t a, b;
};

static std::vector<Shadow> linesweep( const std::vector<Seed>& S, std::vector<Shadow> sh, int64_t uiRStart,
                                      double fAngle ) // harmonization.cpp:182-249
{
    std::sort( sh.begin( ), sh.end( ), []( Shadow xA, Shadow xB ) {
        if( xA.a == xB.a )
            return xA.b > xB.b;
        return xA.a < xB.a;
    } );
    std::vector<Shadow> ends;
    uint64_t x = 0;
    for( auto& t : sh )
    {
        if( x < t.b )
        {
            ends.push_back( t );
            x = t.b;
        }
        else
        {
            double fDistance = deltaDistance( S[ t.seed ], fAngle, uiRStart );
            uint64_t uiPos = ends.size( );
            bool bCloser = true;
            while( uiPos > 0 && ends[ uiPos - 1 ].b >= t.b )
            {
                double fOther = deltaDistance( S[ ends[ uiPos - 1 ].seed ], fAngle, uiRStart );
                if( fOther <= fDistance )
                {
                    bCloser = false;
                    break;
                }
                --uiPos;
            }
            if( bCloser )
            {
                while( !ends.empty( ) && ends.back( ).b >= t.b )
                    ends.pop_back( );
                ends.push_back( t );
            }
        }
    }
    return ends;
}

static std::vector<Seed> harmonizeOne( std::vector<Seed>& in ) // harmonization.cpp:251-373
{
    std::vector<Seed> out;
    if( in.size( ) > 1 )
    {
        std::vector<double> vX, vY;
        for( const auto& s : in )
        {
            vX.push_back( (double)s.r + s.len / 2.0 ), vY.push_back( (double)s.q + s.len / 2.0 );
            vX.push_back( (double)s.r ), vY.push_back( (double)s.q );
            vX.push_back( (double)s.r + s.len ), vY.push_back( (double)s.q + s.len );
        }
        double fMAD = medianAbsoluteDeviation( vY );
        auto si = run_ransac( vX, vY, fMAD );
        in.erase( std::remove_if( in.begin( ), in.end( ),
                                  [ & ]( const Seed& s ) { return deltaDistance( s, si.first, (int64_t)si.second ) > fMAD; } ),
                  in.end( ) );
        std::vector<Shadow> sh;
        for( size_t i = 0; i < in.size( ); i++ )
            sh.push_back( Shadow{ i, (uint64_t)in[ i ].q, (uint64_t)in[ i ].end_ref( ) } );
        auto sh2 = linesweep( in, sh, (int64_t)si.second, si.first );
        sh.clear( );
        for( auto& t : sh2 )
            sh.push_back( Shadow{ t.seed, (uint64_t)in[ t.seed ].r, (uint64_t)in[ t.seed ].end( ) } );
        sh = linesweep( in, sh, (int64_t)si.second, si.first );
        for( auto& t : sh )
            out.push_back( in[ t.seed ] );
        std::sort( out.begin( ), out.end( ), []( const Seed& a, const Seed& b ) {
            if( a.r == b.r )
                return a.q < b.q;
            return a.r < b.r;
        } );
        if( out.size( ) <= 1 )
        {
            out.clear( );
            out.push_back( in[ in.size( ) / 2 ] );
        }
    }
    else if( !in.empty( ) )
        out.push_back( in.front( ) );
    return out;
}

static std::vector<Seed> applyFilters( const Params& P, std::vector<Seed>& in ) // harmonization.cpp:14-173
{
    if( P.gap_cost_cutting )
    {
        int64_t iScore = P.match * in.front( ).len;
        uint64_t uiMaxScore = iScore;
        size_t lastStart = 0, optStart = 0, optEnd = 0;
        for( size_t i = 1; i < in.size( ); i++ )
        {
            iScore += P.match * in[ i ].len;
            uint64_t uiGap = 0;
            if( in[ i ].q > in[ i - 1 ].q )
                uiGap = in[ i ].q - in[ i - 1 ].q;
            if( in[ i ].r > in[ i - 1 ].r )
            {
                if( (uint64_t)( in[ i ].r - in[ i - 1 ].r ) < uiGap )
                {
                    uiGap -= in[ i ].r - in[ i - 1 ].r;
                    if( P.optimistic_gap_estimation )
                        iScore += P.match * ( in[ i ].r - in[ i - 1 ].r );
                }
                else
                {
                    if( P.optimistic_gap_estimation )
                        iScore += P.match * uiGap;
                    uiGap = ( in[ i ].r - in[ i - 1 ].r ) - uiGap;
                }
            }
            uiGap *= P.extend;
            if( uiGap > 0 )
                uiGap += P.gap;
            if( uiGap > (uint64_t)P.sv_penalty && P.sv_penalty != 0 )
                uiGap = (uint64_t)P.sv_penalty;
            if( iScore < (int64_t)uiGap )
            {
                iScore = 0;
                lastStart = i;
            }
            else
                iScore -= uiGap;
            if( iScore > (int64_t)uiMaxScore )
            {
                uiMaxScore = iScore;
                optStart = lastStart;
                optEnd = i;
            }
        }
        // keep [optStart, optEnd]; the remainder stays in `in` for the caller's loop — see :119-132: what is
        // erased from pIn is everything OUTSIDE the optimal run, and then the pointers are swapped.
        if( optEnd != in.size( ) )
            if( ++optEnd != in.size( ) )
                in.erase( in.begin( ) + optEnd, in.end( ) );
        if( optStart != in.size( ) )
            in.erase( in.begin( ), in.begin( ) + optStart );
    }
    std::vector<Seed> ret;
    ret.swap( in ); // pRet.swap(pIn): in becomes empty
    if( ret.size( ) > 2 )
    {
        size_t pre = 0, center = 1;
        while( center < ret.size( ) - 1 )
        {
            Seed& rPre = ret[ pre ];
            Seed& rC = ret[ center ];
            Seed& rPost = ret[ center + 1 ];
            int64_t dPre = rPre.r - (int64_t)rPre.q, dC = rC.r - (int64_t)rC.q, dPost = rPost.r - (int64_t)rPost.q;
            int64_t toPre = std::abs( dPre - dC ), toPost = std::abs( dPost - dC );
            double diff = std::abs( toPre - toPost ) * 2 / ( (double)toPre + toPost );
            if( diff < P.max_delta_dist && (uint64_t)toPre > (uint64_t)P.min_delta_dist )
            {
                rC.len = 0;
                center++;
            }
            else
            {
                center++;
                pre = center - 1;
            }
        }
    }
    return ret;
}

std::vector<SeedSet> harmonization( const Index& I, const Params& P, SoCQueue& Q, int64_t qlen )
{ // harmonization.cpp:374-555
    unsigned uiNumTries = 0;
    uint64_t uiLastHarmScore = 0, uiBestSoCScore = 0;
    unsigned uiSoCRepeatCounter = 0;
    std::vector<SeedSet> out;
    const bool bDoHeuristics = !P.disable_heuristics;
    const unsigned uiMaxTries = P.max_num_soc, uiMinTries = P.min_num_soc;
    const uint64_t uiSwitchQLen = P.switch_qlen;
    while( !Q.maxima.empty( ) )
    {
        if( ++uiNumTries > uiMaxTries )
            break;
        unsigned socIndex;
        auto seedsIn = soc_pop( Q, &socIndex );
        uint64_t uiCurrSoCScore = 0;
        for( auto& s : seedsIn )
            uiCurrSoCScore += s.len;
        if( bDoHeuristics && uiNumTries > uiMinTries )
        {
            if( (uint64_t)qlen > uiSwitchQLen && uiSwitchQLen != 0 )
                if( uiLastHarmScore > uiCurrSoCScore )
                    continue;
            if( uiBestSoCScore * P.soc_score_drop > uiCurrSoCScore && P.soc_score_drop > 0 )
                break;
        }
        uiBestSoCScore = std::max( uiBestSoCScore, uiCurrSoCScore );
        // extractStrand(false) (seed.h:411-426) + unfold
        std::vector<Seed> rev;
        {
            std::vector<Seed> keep;
            for( auto& s : seedsIn )
                ( s.fw == false ? rev : keep ).push_back( s );
            seedsIn.swap( keep );
        }
        for( auto& s : rev )
            s.r = I.ref_len - s.r - 1;
        auto forw = harmonizeOne( seedsIn );
        auto revh = harmonizeOne( rev );
        uint64_t uiCurrHarmScore = 0;
        for( auto& s : forw )
            uiCurrHarmScore += s.len;
        for( auto& s : revh )
            uiCurrHarmScore += s.len;
        if( bDoHeuristics && uiNumTries > uiMinTries )
            if( uiCurrHarmScore < (uint64_t)P.harm_score_min )
                continue;
        if( bDoHeuristics )
            if( uiCurrHarmScore < qlen * P.harm_score_min_rel )
                continue;
        if( bDoHeuristics && uiNumTries > uiMinTries && (uint64_t)qlen > uiSwitchQLen && uiSwitchQLen != 0 )
            if( uiLastHarmScore > uiCurrHarmScore )
                continue;
        while( !forw.empty( ) )
        {
            uiSoCRepeatCounter++;
            out.push_back( SeedSet{ applyFilters( P, forw ), socIndex } );
        }
        while( !revh.empty( ) )
        {
            uiSoCRepeatCounter++;
            // sic: extractStrand() returns a fresh Seeds whose xStats were never copied from the popped SoC, so every
            // reverse-strand set carries index_of_strip == 0 (seed.h:411-426, harmonization.cpp:437, :255)
            out.push_back( SeedSet{ applyFilters( P, revh ), 0 } );
        }
        if( bDoHeuristics && uiNumTries > uiMinTries && (uint64_t)qlen < uiSwitchQLen && uiSwitchQLen != 0 )
        {
            if( !( uiCurrHarmScore + ( qlen * P.score_diff_tolerance ) >= uiLastHarmScore &&
                   uiCurrHarmScore - ( qlen * P.score_diff_tolerance ) <= uiLastHarmScore ) )
                uiSoCRepeatCounter = 0;
            if( uiSoCRepeatCounter >= (unsigned)P.max_score_lookahead && P.max_score_lookahead != 0 )
                break;
        }
        else
            uiSoCRepeatCounter = 0;
        uiLastHarmScore = uiCurrHarmScore;
    }
    if( bDoHeuristics )
        for( unsigned ui = 0; ui < uiSoCRepeatCounter && out.size( ) > uiMinTries; ui++ )
            out.pop_back( );
    return out;
}

} // namespace oracle
