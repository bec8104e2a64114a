// Strip of Consideration + Harmonization for ONE read, host+device, on raw arrays (no allocation).
//
// Replaces, bit-exactly (see SURVEY.md Appendix A for the order-observable details that are reproduced on purpose):
//   StripOfConsiderationSeeds::execute   libs/ma/src/module/stripOfConsideration.cpp:12-161
//   SoCPriorityQueue (push_back_no_overlap, make_heap, rectangularSoC, pop)  libs/ma/inc/ma/container/soc.h:196-419
//   Harmonization::execute / harmonizeOne / linesweep / applyFilters         libs/ma/src/module/harmonization.cpp:14-555
//   run_ransac / RANSAC::computeModel / SACModelLine / lin_regres / MAD      libs/ma/src/sample_consensus/*,
//                                                                            libs/ma/inc/ma/sample_consensus/*
// Sorting and heap operations go through stl_exact.cuh (same element moves as libstdc++); rand() is glibc's TYPE_3
// additive-feedback generator re-implemented below; floating point is IEEE double in the reference's operation order.
#pragma once
#include "fmindex.cuh"
#include <math.h>

namespace ma
{

struct DSeed // 32 bytes
{
    int q, len; // len == 0 marks a seed removed by the artifact filter
    long long r;
    unsigned int amb;
    int fw;
    long long delta;
};

struct DSoCOrder
{
    unsigned long long acc_len;
    unsigned int amb, count;
};
MA_HD inline void soc_add( DSoCOrder& o, const DSeed& s )
{
    o.amb += s.amb, o.count++, o.acc_len += (unsigned long long)s.len;
}
MA_HD inline void soc_sub( DSoCOrder& o, const DSeed& s )
{
    o.amb -= s.amb, o.acc_len -= (unsigned long long)s.len, o.count--;
}
MA_HD inline bool soc_less( const DSoCOrder& a, const DSoCOrder& b ) // soc.h:73-78
{
    if( a.acc_len == b.acc_len )
        return a.amb > b.amb;
    return a.acc_len < b.acc_len;
}
struct DSoC
{
    DSoCOrder o;
    int begin, end;
};

struct HarmParams
{
    int match, gap, extend, sv_penalty;
    int max_num_soc, min_num_soc, soc_width, rectangular_soc;
    double soc_score_drop;
    int harm_score_min;
    double harm_score_min_rel, score_diff_tolerance;
    int max_score_lookahead, switch_qlen;
    double max_delta_dist;
    int min_delta_dist, optimistic_gap_estimation, gap_cost_cutting, disable_heuristics;
    long long genome_size_disable;
};

// glibc rand()/srand(): TYPE_3 (x^31 + x^3 + 1) additive feedback generator, stdlib/random_r.c
struct GlibcRand
{
    int st[ 31 ];
    int f, r;
    MA_HD void seed( unsigned int s )
    {
        if( s == 0 )
            s = 1;
        int word = (int)s;
        st[ 0 ] = word;
        for( int i = 1; i < 31; ++i )
        {
            const long long hi = word / 127773, lo = word % 127773;
            long long w = 16807 * lo - 2836 * hi;
            if( w < 0 )
                w += 2147483647;
            word = (int)w;
            st[ i ] = word;
        }
        f = 3, r = 0;
        for( int k = 0; k < 310; k++ )
            next( );
    }
    MA_HD int next( )
    {
        const unsigned int val = (unsigned int)st[ f ] + (unsigned int)st[ r ];
        st[ f ] = (int)val;
        const int result = (int)( val >> 1 );
        ++f;
        if( f >= 31 )
        {
            f = 0;
            ++r;
        }
        else
        {
            ++r;
            if( r >= 31 )
                r = 0;
        }
        return result;
    }
};

struct Shadow
{
    int seed;
    unsigned long long a, b;
};

// All arrays are caller-provided; n = number of seeds of the read. Capacities in elements:
struct HarmScratch
{
    DSoC* maxima; // n
    long long* vref; // 2n
    DSeed* popF; // n   (popped SoC, forward part; compacted in place by the outlier filter)
    DSeed* popR; // n
    DSeed* outF; // n   (harmonized sets)
    DSeed* outR; // n
    double* X; // 3n
    double* Y; // 3n
    double* tmpA; // 3n
    double* tmpB; // 3n
    int* inl; // 3n
    int* best; // 3n
    Shadow* shA; // n
    Shadow* shB; // n
};
MA_HD inline size_t harm_scratch_bytes( size_t n )
{
    return n * ( sizeof( DSoC ) + 16 + 4 * sizeof( DSeed ) + 2 * sizeof( Shadow ) ) + 3 * n * ( 4 * 8 + 2 * 4 ) + 256;
}
MA_HD inline HarmScratch harm_scratch_carve( unsigned char* p, size_t n )
{
    HarmScratch s;
    auto take = [ & ]( size_t bytes ) {
        unsigned char* r = p;
        p += ( bytes + 15 ) & ~(size_t)15;
        return r;
    };
    s.X = (double*)take( 3 * n * 8 );
    s.Y = (double*)take( 3 * n * 8 );
    s.tmpA = (double*)take( 3 * n * 8 );
    s.tmpB = (double*)take( 3 * n * 8 );
    s.vref = (long long*)take( 2 * n * 8 );
    s.maxima = (DSoC*)take( n * sizeof( DSoC ) );
    s.popF = (DSeed*)take( n * sizeof( DSeed ) );
    s.popR = (DSeed*)take( n * sizeof( DSeed ) );
    s.outF = (DSeed*)take( n * sizeof( DSeed ) );
    s.outR = (DSeed*)take( n * sizeof( DSeed ) );
    s.shA = (Shadow*)take( n * sizeof( Shadow ) );
    s.shB = (Shadow*)take( n * sizeof( Shadow ) );
    s.inl = (int*)take( 3 * n * 4 );
    s.best = (int*)take( 3 * n * 4 );
    return s;
}
// upper bound of harm_scratch_carve's consumption
MA_HD inline size_t harm_scratch_need( size_t n )
{
    return 4 * ( ( 3 * n * 8 + 15 ) & ~(size_t)15 ) + ( ( 2 * n * 8 + 15 ) & ~(size_t)15 ) +
           ( ( n * sizeof( DSoC ) + 15 ) & ~(size_t)15 ) + 4 * ( ( n * sizeof( DSeed ) + 15 ) & ~(size_t)15 ) +
           2 * ( ( n * sizeof( Shadow ) + 15 ) & ~(size_t)15 ) + 2 * ( ( 3 * n * 4 + 15 ) & ~(size_t)15 );
}

// ------------------------------------------------------------------------------------------------ SoC
MA_HD inline void soc_adjust( const DSeed* S, DSoCOrder& o, int cutS, int cutE, int cntS, int cntE ) // soc.h:338-356
{
    if( cutE - cutS <= cntE - cntS )
        for( int i = cutS; i < cutE; i++ )
            soc_sub( o, S[ i ] );
    else
    {
        DSoCOrder n{ 0, 0, 0 };
        for( int i = cntS; i < cntE; i++ )
            soc_add( n, S[ i ] );
        o = n;
    }
}

// builds the SoC queue over S[0..n) (S is re-ordered); returns the number of windows in maxima[]
MA_HD inline int soc_build( const DevIndex& I, const HarmParams& P, DSeed* S, int n, int qlen, DSoC* maxima,
                            long long* vref )
{
    if( n == 0 )
        return 0;
    double fMinLen = (double)P.harm_score_min_rel * qlen;
    if( fMinLen < (double)P.harm_score_min )
        fMinLen = (double)P.harm_score_min;
    if( (unsigned long long)P.genome_size_disable >= (unsigned long long)( 2 * I.fwd_len ) )
        fMinLen = 0;
    const unsigned long long uiMin = (unsigned long long)fMinLen;
    const unsigned long long strip = P.soc_width != 0
                                         ? (unsigned long long)P.soc_width
                                         : (unsigned long long)( ( (long long)P.match * qlen - P.gap ) / P.extend );
    stl::sort( S, S + n, []( const DSeed& a, const DSeed& b ) { return a.delta < b.delta; } );
    int nMax = 0;
    DSoCOrder cur{ 0, 0, 0 };
    int s = 0, e = 0;
    long long cidS = seq_id_for_position( I, S[ 0 ].r ), cidE = cidS;
    auto inContig = [ & ]( long long id, long long pos ) {
        return I.contig_start[ id ] <= pos && pos < I.contig_start[ id ] + I.contig_len[ id ];
    };
    while( e != n && s != n )
    {
        while( !inContig( cidS, S[ s ].r ) )
            cidS += 1;
        while( e != n && (unsigned long long)S[ s ].delta + strip >= (unsigned long long)S[ e ].delta &&
               cidS == cidE )
        {
            soc_add( cur, S[ e ] );
            e++;
            if( e != n )
                while( !inContig( cidE, S[ e ].r ) )
                    cidE += 1;
        }
        if( (double)cur.acc_len >= fMinLen )
        { // push_back_no_overlap (soc.h:362-404)
            DSoCOrder c = cur;
            int itS = s;
            const int itE = e;
            bool bPush = true;
            while( nMax > 0 && maxima[ nMax - 1 ].end > itS )
            {
                DSoC& b = maxima[ nMax - 1 ];
                if( soc_less( b.o, c ) )
                {
                    soc_adjust( S, b.o, itS, b.end, b.begin, itS );
                    b.end = itS;
                    if( b.o.acc_len < uiMin || b.o.acc_len == 0 )
                        nMax--;
                }
                else
                {
                    soc_adjust( S, c, itS, b.end, b.end, itE );
                    itS = b.end;
                    if( c.acc_len < uiMin || c.acc_len == 0 )
                    {
                        bPush = false;
                        break;
                    }
                }
            }
            if( bPush )
                maxima[ nMax++ ] = DSoC{ c, itS, itE };
        }
        soc_sub( cur, S[ s ] );
        s++;
    }
    auto heapOrder = []( const DSoC& a, const DSoC& b ) { return soc_less( a.o, b.o ); };
    stl::make_heap( maxima, (long)nMax, heapOrder );
    if( P.rectangular_soc )
    { // soc.h:196-231: windows are rebuilt over start_ref order, the heap array is NOT re-heapified
        for( int m = 0; m < nMax; m++ )
        {
            long long lo = S[ maxima[ m ].begin ].r, hi = lo;
            for( int i = maxima[ m ].begin; i != maxima[ m ].end; i++ )
            {
                lo = S[ i ].r < lo ? S[ i ].r : lo;
                hi = S[ i ].r > hi ? S[ i ].r : hi;
            }
            vref[ 2 * m ] = lo, vref[ 2 * m + 1 ] = hi;
        }
        stl::sort( S, S + n, []( const DSeed& a, const DSeed& b ) { return a.r < b.r; } );
        for( int m = 0; m < nMax; m++ )
        {
            int lo = 0, cnt = n; // std::lower_bound
            while( cnt > 0 )
            {
                const int half = cnt >> 1;
                if( S[ lo + half ].r < vref[ 2 * m ] )
                    lo = lo + half + 1, cnt = cnt - half - 1;
                else
                    cnt = half;
            }
            DSoC w{ { 0, 0, 0 }, lo, lo };
            int it = lo;
            while( it != n && S[ it ].r <= vref[ 2 * m + 1 ] )
                soc_add( w.o, S[ it ] ), it++;
            w.end = it;
            maxima[ m ] = w;
        }
    }
    return nMax;
}

// ------------------------------------------------------------------------------------------------ RANSAC
MA_HD inline double median_of( double* a, int n ) // test_ransac.h:20-39 (a is sorted in place)
{
    stl::sort( a, a + n, []( double x, double y ) { return x < y; } );
    if( n == 0 )
        return 0;
    if( n == 1 )
        return a[ 0 ];
    if( n % 2 == 0 )
        return ( a[ n / 2 - 1 ] + a[ n / 2 ] ) / 2;
    return a[ n / 2 ];
}

MA_HD inline double ma_nan( )
{
#if defined( __CUDA_ARCH__ )
    return __longlong_as_double( 0x7ff8000000000000ll );
#else
    return __builtin_nan( "" );
#endif
}

// returns (angle, intercept) like run_ransac (test_ransac.cpp:8-100); N points in X, Y
MA_HD inline void run_ransac( const double* X, const double* Y, int N, double fMAD, GlibcRand& rng, int* inl,
                              int* best, double* dx, double* dy, double& angle, double& icpt )
{
    int iterations = 0, n_best = -2147483647, nBestInl = 0;
    double k = 1.0;
    bool bHave = false;
    const double threshold = fMAD, probability = 0.99;
    const double eps = 2.220446049250313e-16;
    const double pi = 3.14159265358979323846; // std::acos(-1)
    while( iterations < k )
    {
        int s0, s1;
        { // SACModelLine::getSamples (sac_model_line.cpp:49-76)
            const double trand = N / ( 2147483647 + 1.0 );
            int idx = (int)( rng.next( ) * trand );
            s0 = idx;
            int iter = 0;
            do
            {
                idx = (int)( rng.next( ) * trand );
                s1 = idx;
                iter++;
                if( iter > 1000 )
                    break;
                iterations++;
            } while( s1 == s0 );
            iterations--;
        }
        const double m0 = X[ s0 ], m1 = Y[ s0 ], m3 = X[ s1 ], m4 = Y[ s1 ];
        double dH = m0 - m3, dV = m1 - m4;
        if( dH <= 0 && dV <= 0 )
            dH *= -1, dV *= -1;
        double dAngle = -90;
        if( dH > 0 && dV > 0 )
            dAngle = atan( dV / dH ) * 180 / pi;
        if( dAngle >= 20 && dAngle <= 70 )
        {
            const double sqr_threshold = threshold * threshold;
            int nInl = 0;
            const double p3x = m3 - m0, p3y = m4 - m1, p3z = 0.0 - 0.0;
            for( int i = 0; i < N; i++ )
            { // 3-D cross product form with z = 0 (sac_model_line.cpp:86-131, sac_model.h:70-78)
                const double p4x = m3 - X[ i ], p4y = m4 - Y[ i ], p4z = 0.0 - 0.0;
                const double cx = p4y * p3z - p4z * p3y, cy = p4z * p3x - p4x * p3z, cz = p4x * p3y - p4y * p3x;
                const double sqr_distance = ( cx * cx + cy * cy + cz * cz ) / ( p3x * p3x + p3y * p3y + p3z * p3z );
                if( sqr_distance < sqr_threshold )
                    inl[ nInl++ ] = i;
            }
            if( nInl > n_best )
            {
                n_best = nInl;
                for( int i = 0; i < nInl; i++ )
                    best[ i ] = inl[ i ];
                nBestInl = nInl;
                bHave = true;
                const double w = (double)nInl / (double)N;
                double p_no = 1 - w * w; // pow(w, 2.0)
                p_no = eps > p_no ? eps : p_no;
                p_no = ( 1 - eps ) < p_no ? ( 1 - eps ) : p_no;
                k = log( 1 - probability ) / log( p_no );
            }
        }
        else
            continue;
        iterations += 1;
        if( iterations > 100 )
            break;
    }
    if( !bHave )
    {
        angle = ma_nan( ), icpt = ma_nan( );
        return;
    }
    // lin_regres over the inliers (lin_regres.h:58-147)
    const int n = nBestInl;
    double sum = 0;
    for( int i = 0; i < n; i++ )
        sum = sum + X[ best[ i ] ];
    const double mean_x = sum / (double)n;
    sum = 0;
    for( int i = 0; i < n; i++ )
        sum = sum + Y[ best[ i ] ];
    const double mean_y = sum / (double)n;
    double sx = 0;
    for( int i = 0; i < n; i++ )
    {
        dx[ i ] = X[ best[ i ] ] - mean_x;
        sx = sx + dx[ i ] * dx[ i ];
    }
    for( int i = 0; i < n; i++ )
        dy[ i ] = Y[ best[ i ] ] - mean_y;
    double sum_xy = 0;
    for( int i = 0; i < n; i++ )
        sum_xy = sum_xy + dx[ i ] * dy[ i ];
    const double slope = sum_xy / sx;
    const double intercept = mean_y - slope * mean_x;
    angle = atan( slope );
    icpt = -intercept / slope;
}

// ------------------------------------------------------------------------------------------------ Harmonization
#define MA_HARM_PI 3.14159265
MA_HD inline double delta_distance( const DSeed& s, double fAngle, long long uiRStart ) // harmonization.h:82-89
{
    const double y = s.r + s.q / tan( MA_HARM_PI / 2 - fAngle );
    const double x = ( y - uiRStart ) * sin( fAngle );
    const double x_1 = s.q / sin( MA_HARM_PI / 2 - fAngle );
    return fabs( x - x_1 );
}

// harmonization.cpp:182-249: sh[0..n) -> ends[0..return)
MA_HD MA_NOINLINE inline int linesweep( const DSeed* S, Shadow* sh, int n, Shadow* ends, long long uiRStart, double fAngle )
{
    stl::sort( sh, sh + n, []( const Shadow& xA, const Shadow& xB ) {
        if( xA.a == xB.a )
            return xA.b > xB.b;
        return xA.a < xB.a;
    } );
    int nE = 0;
    unsigned long long x = 0;
    for( int i = 0; i < n; i++ )
    {
        const Shadow t = sh[ i ];
        if( x < t.b )
        {
            ends[ nE++ ] = t;
            x = t.b;
        }
        else
        {
            const double fDistance = delta_distance( S[ t.seed ], fAngle, uiRStart );
            int uiPos = nE;
            bool bCloser = true;
            while( uiPos > 0 && ends[ uiPos - 1 ].b >= t.b )
            {
                const double fOther = delta_distance( S[ ends[ uiPos - 1 ].seed ], fAngle, uiRStart );
                if( fOther <= fDistance )
                {
                    bCloser = false;
                    break;
                }
                --uiPos;
            }
            if( bCloser )
            {
                while( nE > 0 && ends[ nE - 1 ].b >= t.b )
                    nE--;
                ends[ nE++ ] = t;
            }
        }
    }
    return nE;
}

MA_HD inline long long double_to_ll( double d ) // (int64_t)d as x86-64 cvttsd2si does it
{
    if( !( d == d ) || d >= 9223372036854775808.0 || d < -9223372036854775808.0 )
        return (long long)0x8000000000000000ull;
    return (long long)d;
}

// harmonization.cpp:251-373. in[0..nIn) is compacted in place by the outlier filter; out receives the result.
// (not inlined: two call sites, and the instruction footprint of socharm_kernel is what bounds it)
MA_HD MA_NOINLINE inline int harmonize_one( DSeed* in, int& nIn, DSeed* out, HarmScratch& W, GlibcRand& rng )
{
    int nOut = 0;
    if( nIn > 1 )
    {
        const int N = 3 * nIn;
        for( int i = 0; i < nIn; i++ )
        {
            const DSeed& s = in[ i ];
            W.X[ 3 * i ] = (double)s.r + s.len / 2.0, W.Y[ 3 * i ] = (double)s.q + s.len / 2.0;
            W.X[ 3 * i + 1 ] = (double)s.r, W.Y[ 3 * i + 1 ] = (double)s.q;
            W.X[ 3 * i + 2 ] = (double)s.r + s.len, W.Y[ 3 * i + 2 ] = (double)s.q + s.len;
        }
        // medianAbsoluteDeviation (test_ransac.h:57-74)
        for( int i = 0; i < N; i++ )
            W.tmpA[ i ] = W.Y[ i ];
        const double median = median_of( W.tmpA, N );
        for( int i = 0; i < N; i++ )
            W.tmpB[ i ] = W.Y[ i ] - median < 0 ? -( W.Y[ i ] - median ) : W.Y[ i ] - median;
        const double fMAD = median_of( W.tmpB, N );
        double angle, icpt;
        run_ransac( W.X, W.Y, N, fMAD, rng, W.inl, W.best, W.tmpA, W.tmpB, angle, icpt );
        const long long rStart = double_to_ll( icpt );
        int k = 0;
        for( int i = 0; i < nIn; i++ ) // erase(remove_if(deltaDistance > MAD))
            if( !( delta_distance( in[ i ], angle, rStart ) > fMAD ) )
                in[ k++ ] = in[ i ];
        nIn = k;
        for( int i = 0; i < nIn; i++ )
            W.shA[ i ] = Shadow{ i, (unsigned long long)in[ i ].q, (unsigned long long)( in[ i ].r + in[ i ].len ) };
        int n2 = linesweep( in, W.shA, nIn, W.shB, rStart, angle );
        for( int i = 0; i < n2; i++ )
        {
            const int sd = W.shB[ i ].seed;
            W.shA[ i ] = Shadow{ sd, (unsigned long long)in[ sd ].r, (unsigned long long)( in[ sd ].q + in[ sd ].len ) };
        }
        n2 = linesweep( in, W.shA, n2, W.shB, rStart, angle );
        for( int i = 0; i < n2; i++ )
            out[ nOut++ ] = in[ W.shB[ i ].seed ];
        stl::sort( out, out + nOut, []( const DSeed& a, const DSeed& b ) {
            if( a.r == b.r )
                return a.q < b.q;
            return a.r < b.r;
        } );
        if( nOut <= 1 )
        {
            nOut = 0;
            out[ nOut++ ] = in[ nIn / 2 ]; // sic: reads in[0] of an emptied vector if the filter removed everything
        }
    }
    else if( nIn != 0 )
        out[ nOut++ ] = in[ 0 ];
    return nOut;
}

// harmonization.cpp:14-173. set[0..n) -> kept run moved to the front; returns its length (the caller's vector
// becomes empty afterwards, pRet.swap(pIn))
MA_HD inline int apply_filters( const HarmParams& P, DSeed* in, int n )
{
    int b = 0, e = n; // kept range [b, e)
    if( P.gap_cost_cutting )
    {
        long long iScore = (long long)P.match * in[ 0 ].len;
        unsigned long long uiMaxScore = (unsigned long long)iScore;
        int lastStart = 0, optStart = 0, optEnd = 0;
        for( int i = 1; i < n; i++ )
        {
            iScore += (long long)P.match * in[ i ].len;
            unsigned long long uiGap = 0;
            if( in[ i ].q > in[ i - 1 ].q )
                uiGap = (unsigned long long)( in[ i ].q - in[ i - 1 ].q );
            if( in[ i ].r > in[ i - 1 ].r )
            {
                const unsigned long long dr = (unsigned long long)( in[ i ].r - in[ i - 1 ].r );
                if( dr < uiGap )
                {
                    uiGap -= dr;
                    if( P.optimistic_gap_estimation )
                        iScore += (long long)( (unsigned long long)P.match * dr );
                }
                else
                {
                    if( P.optimistic_gap_estimation )
                        iScore += (long long)( (unsigned long long)P.match * uiGap );
                    uiGap = dr - uiGap;
                }
            }
            uiGap *= (unsigned long long)P.extend;
            if( uiGap > 0 )
                uiGap += (unsigned long long)P.gap;
            if( uiGap > (unsigned long long)P.sv_penalty && P.sv_penalty != 0 )
                uiGap = (unsigned long long)P.sv_penalty;
            if( iScore < (long long)uiGap )
            {
                iScore = 0;
                lastStart = i;
            }
            else
                iScore -= (long long)uiGap;
            if( iScore > (long long)uiMaxScore )
            {
                uiMaxScore = (unsigned long long)iScore;
                optStart = lastStart;
                optEnd = i;
            }
        }
        b = optStart;
        e = optEnd + 1; // n >= 1 so optEnd != end(); ++optEnd, erase [optEnd, end)
    }
    const int m = e - b;
    if( b > 0 )
        for( int i = 0; i < m; i++ )
            in[ i ] = in[ b + i ];
    if( m > 2 )
    {
        int pre = 0, center = 1;
        while( center < m - 1 )
        {
            DSeed& rPre = in[ pre ];
            DSeed& rC = in[ center ];
            DSeed& rPost = in[ center + 1 ];
            const long long dPre = rPre.r - (long long)rPre.q, dC = rC.r - (long long)rC.q,
                            dPost = rPost.r - (long long)rPost.q;
            const long long toPre = dPre - dC < 0 ? -( dPre - dC ) : dPre - dC,
                            toPost = dPost - dC < 0 ? -( dPost - dC ) : dPost - dC;
            const long long ad = toPre - toPost < 0 ? -( toPre - toPost ) : toPre - toPost;
            const double diff = ad * 2 / ( (double)toPre + toPost );
            if( diff < P.max_delta_dist && (unsigned long long)toPre > (unsigned long long)P.min_delta_dist )
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
    return m;
}

// Sink: void set( const DSeed* seeds, int n, unsigned soc_index )
// Harmonization::execute over the SoC queue that soc_build left in W.maxima[0..nMax) (S: the read's seeds as soc_build
// ordered them): pops, RANSAC / linesweep / filters per strip, the heuristics that end the loop
template <class Sink>
MA_HD inline void soc_harm_pops( const DevIndex& I, const HarmParams& P, const DSeed* S, int n, int qlen,
                                 unsigned int srand_seed, HarmScratch& W, Sink& sink, int nMax )
{
    auto heapOrder = []( const DSoC& a, const DSoC& b ) { return soc_less( a.o, b.o ); };
    GlibcRand rng;
    rng.seed( srand_seed );
    unsigned int uiNumTries = 0, uiSoCRepeatCounter = 0, nextIndex = 0;
    unsigned long long uiLastHarmScore = 0, uiBestSoCScore = 0;
    const bool bDoHeuristics = !P.disable_heuristics;
    const unsigned int uiMaxTries = (unsigned)P.max_num_soc, uiMinTries = (unsigned)P.min_num_soc;
    const unsigned long long uiSwitchQLen = (unsigned long long)P.switch_qlen;
    while( nMax > 0 )
    {
        if( ++uiNumTries > uiMaxTries )
            break;
        // SoCPriorityQueue::pop (soc.h:240-284) + extractStrand(false) + un-folding of reverse strand seeds
        const DSoC top = W.maxima[ 0 ];
        const unsigned int socIndex = nextIndex++;
        int nF = 0, nR = 0;
        unsigned long long uiCurrSoCScore = 0;
        for( int i = top.begin; i != n && i != top.end; i++ )
        {
            uiCurrSoCScore += (unsigned long long)S[ i ].len;
            if( S[ i ].fw )
                W.popF[ nF++ ] = S[ i ];
            else
            {
                W.popR[ nR ] = S[ i ];
                W.popR[ nR ].r = I.ref_len - S[ i ].r - 1;
                nR++;
            }
        }
        stl::pop_heap( W.maxima, (long)nMax, heapOrder );
        nMax--;
        if( bDoHeuristics && uiNumTries > uiMinTries )
        {
            if( (unsigned long long)qlen > uiSwitchQLen && uiSwitchQLen != 0 )
                if( uiLastHarmScore > uiCurrSoCScore )
                    continue;
            if( uiBestSoCScore * P.soc_score_drop > uiCurrSoCScore && P.soc_score_drop > 0 )
                break;
        }
        uiBestSoCScore = uiBestSoCScore > uiCurrSoCScore ? uiBestSoCScore : uiCurrSoCScore;
        int nOF = harmonize_one( W.popF, nF, W.outF, W, rng );
        int nOR = harmonize_one( W.popR, nR, W.outR, W, rng );
        unsigned long long uiCurrHarmScore = 0;
        for( int i = 0; i < nOF; i++ )
            uiCurrHarmScore += (unsigned long long)W.outF[ i ].len;
        for( int i = 0; i < nOR; i++ )
            uiCurrHarmScore += (unsigned long long)W.outR[ i ].len;
        if( bDoHeuristics && uiNumTries > uiMinTries )
            if( uiCurrHarmScore < (unsigned long long)P.harm_score_min )
                continue;
        if( bDoHeuristics )
            if( uiCurrHarmScore < qlen * P.harm_score_min_rel )
                continue;
        if( bDoHeuristics && uiNumTries > uiMinTries && (unsigned long long)qlen > uiSwitchQLen && uiSwitchQLen != 0 )
            if( uiLastHarmScore > uiCurrHarmScore )
                continue;
        if( nOF > 0 )
        {
            uiSoCRepeatCounter++;
            const int m = apply_filters( P, W.outF, nOF );
            sink.set( W.outF, m, socIndex );
        }
        if( nOR > 0 )
        {
            uiSoCRepeatCounter++;
            const int m = apply_filters( P, W.outR, nOR );
            sink.set( W.outR, m, 0 ); // sic: reverse-strand sets carry index_of_strip 0 (see oracle note)
        }
        if( bDoHeuristics && uiNumTries > uiMinTries && (unsigned long long)qlen < uiSwitchQLen && uiSwitchQLen != 0 )
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
        sink.pop_back( uiSoCRepeatCounter, uiMinTries ); // for(ui < counter && size > uiMinTries) pop_back()
}

// StripOfConsiderationSeeds::execute + Harmonization::execute for one read
template <class Sink>
MA_HD inline void soc_harm_read( const DevIndex& I, const HarmParams& P, DSeed* S, int n, int qlen,
                                 unsigned int srand_seed, HarmScratch& W, Sink& sink, int uiMinTriesOutCount )
{
    (void)uiMinTriesOutCount;
    const int nMax = soc_build( I, P, S, n, qlen, W.maxima, W.vref );
    soc_harm_pops( I, P, S, n, qlen, srand_seed, W, sink, nMax );
}

} // namespace ma
