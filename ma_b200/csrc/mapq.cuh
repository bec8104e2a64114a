// MappingQuality and PairedReads on the alignment records of the NW stage.
//
// Replaces, with identical results,
//   libs/ma/src/module/mappingQuality.cpp:11-131   (score sort, supplementary / secondary flags, mapping quality,
//                                                   larger()-sort, n-best and minimal-score filters)
//   libs/ma/src/module/pairedReads.cpp:15-121      (all mate combinations, insert-size bonus, pair quality)
//   libs/ma/inc/ma/container/alignment.h:239-246, 659-742, 819-843 (getNumSeeds, overlap, larger)
// One thread per read / per pair; both std::sorts are the exact libstdc++ emulation of stl_exact.cuh (the order of
// equal scores decides which alignment becomes primary). Host+device: tests/hostsim runs the same code on the CPU.
#pragma once
#include "nwglue.cuh"
#include <cmath>

namespace ma
{

#define MA_ALN_SECONDARY 1
#define MA_ALN_SUPPLEMENTARY 2
#define MA_ALN_FIRST_MATE 4

struct MapqParams
{
    int match;
    int report_n, min_alignment_score, max_supplementary_per_prim;
    double max_overlap_supplementary;
    double paired_mean, paired_std, paired_bonus;
};

MA_HD inline int run_type( unsigned int w )
{
    return (int)( w & 7 );
}
MA_HD inline unsigned long long run_len( unsigned int w )
{
    return (unsigned long long)( w >> 3 );
}

MA_HD inline int aln_num_seeds( const DAln& a, const unsigned int* runs )
{
    int n = 0;
    for( int i = 0; i < a.n_runs; i++ )
        if( run_type( runs[ a.run_off + i ] ) == 0 )
            n++;
    return n;
}

// Alignment::overlap (alignment.h:659-742); match types: seed 0, match 1, missmatch 2, insertion 3, deletion 4
MA_HD inline double aln_overlap( const DAln& A, const DAln& O, const unsigned int* runs )
{
    typedef unsigned long long u64;
    const u64 uiS = (u64)( A.begin_q > O.begin_q ? A.begin_q : O.begin_q ),
              uiE = (u64)( A.end_q < O.end_q ? A.end_q : O.end_q );
    if( uiS >= uiE )
        return 0;
    const unsigned int* da = runs + A.run_off;
    const unsigned int* dob = runs + O.run_off;
    u64 uiOverlap = 0, qa = (u64)A.begin_q, qo = (u64)O.begin_q;
    int i = 0, io = 0;
    while( qa + run_len( da[ i ] ) < uiS )
    {
        if( run_type( da[ i ] ) != 4 )
            qa += run_len( da[ i ] );
        i++;
    }
    while( qo + run_len( dob[ io ] ) < uiS )
    {
        if( run_type( dob[ io ] ) != 4 )
            qo += run_len( dob[ io ] );
        io++;
    }
    while( qa < uiE && qo < uiE && i < A.n_runs && io < O.n_runs )
    {
        const u64 l = run_type( da[ i ] ) != 4 ? run_len( da[ i ] ) : 0;
        const u64 lo = run_type( dob[ io ] ) != 4 ? run_len( dob[ io ] ) : 0;
        u64 s = qa > qo ? qa : qo;
        s = s > uiS ? s : uiS;
        u64 e = qa + l < qo + lo ? qa + l : qo + lo;
        e = e < uiE ? e : uiE;
        const u64 cur = s < e ? e - s : 0;
        if( run_type( da[ i ] ) != 3 && run_type( dob[ io ] ) != 3 )
            uiOverlap += cur;
        if( qa + l < qo + lo )
            qa += l, i++;
        else
            qo += lo, io++;
    }
    const u64 sa = (u64)( A.end_q - A.begin_q ), so = (u64)( O.end_q - O.begin_q );
    return (double)uiOverlap / (double)( sa < so ? sa : so );
}

// MappingQuality::execute for the n alignments al[0..n) of one read (their `rank` is the position in the
// NeedlemanWunsch result). Sets flags, mapq and rank_mq (position in the returned vector, -1 = not reported).
// ord: scratch of n ints. Returns the size of the returned vector.
MA_HD inline int mapping_quality_read( const MapqParams& P, DAln* al, int n, const unsigned int* runs, long long qlen,
                                       int* ord )
{
    for( int i = 0; i < n; i++ )
    {
        ord[ al[ i ].rank ] = i;
        al[ i ].flags = 0, al[ i ].rank_mq = -1, al[ i ].pair_rank = -1;
        al[ i ].mapq = NAN; // Alignment::fMappingQuality starts as NAN (alignment.h:75)
    }
    if( n == 0 )
        return 0;
    stl::sort( ord, ord + n, [ & ]( int a, int b ) { return al[ a ].score > al[ b ].score; } );
    DAln& F = al[ ord[ 0 ] ];
    int nSupp = 0;
    for( int i = 1; i < n; i++ )
    {
        DAln& c = al[ ord[ i ] ];
        c.mapq = 0.0;
        if( nSupp < P.max_supplementary_per_prim && aln_overlap( c, F, runs ) < P.max_overlap_supplementary )
            c.flags = MA_ALN_SUPPLEMENTARY, nSupp++;
        else
            c.flags = MA_ALN_SECONDARY;
    }
    const double dMax = (double)( (unsigned long long)P.match * (unsigned long long)qlen );
    if( n - nSupp >= 2 )
    {
        int k = 1;
        while( al[ ord[ k ] ].flags & MA_ALN_SUPPLEMENTARY )
            k++;
        const long long s1 = F.score, s2 = al[ ord[ k ] ].score;
        F.mapq = s1 == 0 ? 0.0 : (double)( s1 - s2 ) / (double)s1;
    }
    else
        F.mapq = (double)F.score / dMax;
    if( aln_num_seeds( F, runs ) <= 1 )
        F.mapq /= 2;
    if( (double)F.score >= dMax * 0.8 && n >= 3 )
        F.mapq *= 2;
    if( F.mapq > 1 )
        F.mapq = 1;
    if( nSupp > 0 )
    {
        for( int i = 1; i < n; i++ )
            if( al[ ord[ i ] ].flags & MA_ALN_SUPPLEMENTARY )
                al[ ord[ i ] ].mapq = F.mapq;
        stl::sort( ord, ord + n, [ & ]( int a, int b ) { // Alignment::larger
            const int ua = ( al[ a ].flags & MA_ALN_SUPPLEMENTARY ) ? 1 : ( al[ a ].flags & MA_ALN_SECONDARY ) ? 2 : 0;
            const int ub = ( al[ b ].flags & MA_ALN_SUPPLEMENTARY ) ? 1 : ( al[ b ].flags & MA_ALN_SECONDARY ) ? 2 : 0;
            if( ua != ub )
                return ua < ub;
            if( al[ a ].score == al[ b ].score )
                return al[ a ].soc_index < al[ b ].soc_index;
            return al[ a ].score > al[ b ].score;
        } );
    }
    int m = n;
    if( P.report_n != 0 && n > P.report_n + nSupp )
        m = P.report_n + nSupp;
    int out = 0;
    for( int i = 0; i < m; i++ )
        if( !( al[ ord[ i ] ].score < (long long)P.min_alignment_score ) )
            al[ ord[ i ] ].rank_mq = out++;
    return out;
}

// PairedReads::execute for the MappingQuality results of the two mates (a1[0..n1), a2[0..n2) with rank_mq set).
// Marks the returned alignments with pair_rank (position in the returned vector) and updates flags / mapq.
// ord1/ord2: scratch of n1/n2 ints; sc/pi/pj/ps: scratch of (reported1 * reported2) entries. Returns the number of
// returned alignments, or -1 if the scratch is too small.
MA_HD inline int paired_reads_pair( const MapqParams& P, long long ref_len, DAln* a1, int n1, long long qlen1, DAln* a2,
                                    int n2, long long qlen2, const unsigned int* runs, int* ord1, int* ord2,
                                    long long* sc, int* meta, int cap )
{
    int m1 = 0, m2 = 0;
    for( int i = 0; i < n1; i++ )
        if( a1[ i ].rank_mq >= 0 )
            ord1[ a1[ i ].rank_mq ] = i, m1++;
    for( int i = 0; i < n2; i++ )
        if( a2[ i ].rank_mq >= 0 )
            ord2[ a2[ i ].rank_mq ] = i, m2++;
    for( int i = 0; i < m1; i++ )
        a1[ ord1[ i ] ].flags |= MA_ALN_FIRST_MATE;
    if( m1 == 0 )
    {
        for( int i = 0; i < m2; i++ )
            a2[ ord2[ i ] ].pair_rank = i;
        return m2;
    }
    if( m2 == 0 )
    {
        for( int i = 0; i < m1; i++ )
            a1[ ord1[ i ] ].pair_rank = i;
        return m1;
    }
    if( m1 * m2 > cap )
        return -1;
    // candidate k: score sc[k], meta[k] = paired << 30 | i << 15 | j ; idx[] is what std::sort permutes
    int nc = 0;
    const unsigned long long mean = (unsigned long long)P.paired_mean;
    for( int i = 0; i < m1; i++ )
    {
        const DAln& A1 = a1[ ord1[ i ] ];
        if( A1.length == 0 )
            continue;
        for( int j = 0; j < m2; j++ )
        {
            const DAln& A2 = a2[ ord2[ j ] ];
            if( A2.length == 0 )
                continue;
            long long iScore = A1.score + A2.score;
            int paired = 0;
            if( ( A1.begin_ref >= ref_len / 2 ) != ( A2.begin_ref >= ref_len / 2 ) )
            {
                const unsigned long long p1 = (unsigned long long)A1.begin_ref,
                                         p2 = (unsigned long long)ref_len - ( (unsigned long long)A2.begin_ref + 1 );
                const unsigned long long d = p1 < p2 ? p2 - p1 : p1 - p2;
                if( (double)d >= (double)mean - P.paired_std * 3 && (double)d <= (double)mean + P.paired_std * 3 )
                {
                    iScore = (long long)( (double)iScore * P.paired_bonus );
                    paired = 1;
                }
            }
            sc[ nc ] = iScore, meta[ nc ] = paired << 30 | i << 15 | j;
            nc++;
        }
    }
    if( nc == 0 )
        return -1; // the reference reads vScores[ 0 ] of an empty vector here
    int* idx = meta + cap; // second half of the scratch
    for( int k = 0; k < nc; k++ )
        idx[ k ] = k;
    stl::sort( idx, idx + nc, [ & ]( int a, int b ) {
        if( sc[ a ] == sc[ b ] )
            return ( meta[ a ] >> 30 ) && !( meta[ b ] >> 30 );
        return sc[ a ] > sc[ b ];
    } );
    const int k0 = idx[ 0 ];
    DAln& B1 = a1[ ord1[ ( meta[ k0 ] >> 15 ) & 0x7fff ] ];
    DAln& B2 = a2[ ord2[ meta[ k0 ] & 0x7fff ] ];
    B1.flags &= ~( MA_ALN_SECONDARY | MA_ALN_SUPPLEMENTARY );
    B2.flags &= ~( MA_ALN_SECONDARY | MA_ALN_SUPPLEMENTARY );
    if( ( meta[ k0 ] >> 30 ) && nc > 1 )
    {
        float fMapQ = ( (float)( sc[ k0 ] - sc[ idx[ 1 ] ] ) ) / (float)sc[ k0 ];
        if( aln_num_seeds( B1, runs ) <= 1 && aln_num_seeds( B2, runs ) <= 1 )
            fMapQ /= 2;
        if( (double)B1.score >= (double)( (unsigned long long)P.match * (unsigned long long)qlen1 ) * 0.8 && m1 >= 3 )
            fMapQ *= 2;
        else if( (double)B2.score >= (double)( (unsigned long long)P.match * (unsigned long long)qlen2 ) * 0.8 && m2 >= 3 )
            fMapQ *= 2;
        if( fMapQ > 1 )
            fMapQ = 1;
        B1.mapq = fMapQ, B2.mapq = fMapQ;
    }
    B1.pair_rank = 0, B2.pair_rank = 1;
    return 2;
}

} // namespace ma
