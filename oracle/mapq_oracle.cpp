// CPU ORACLE — test infrastructure only (see oracle.h).
// MappingQuality and PairedReads, restated from
//   libs/ma/src/module/mappingQuality.cpp:11-131, libs/ma/inc/ma/module/mappingQuality.h:26-36
//   libs/ma/src/module/pairedReads.cpp:15-121,    libs/ma/inc/ma/module/pairedReads.h:43-55
//   libs/ma/inc/ma/container/alignment.h:239-246 (getNumSeeds), :659-742 (overlap), :819-843 (larger)
//   libs/ma/inc/ma/container/pack.h:900-927 (strand helpers)
// Uses libstdc++'s std::sort like the reference (the order of equal keys is part of the result).
#include "ma_oracle.h"
#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <tuple>

namespace oracle
{

static size_t num_seeds( const Alignment& A )
{
    size_t n = 0;
    for( auto& d : A.data )
        if( d.first == MT_SEED )
            n++;
    return n;
}

// Alignment::overlap (alignment.h:659-742)
static double overlap( const Alignment& A, const Alignment& O )
{
    const uint64_t uiS = (uint64_t)std::max( A.begin_q, O.begin_q ), uiE = (uint64_t)std::min( A.end_q, O.end_q );
    if( uiS >= uiE )
        return 0;
    uint64_t uiOverlap = 0, uiQpos = (uint64_t)A.begin_q, uiQposO = (uint64_t)O.begin_q;
    size_t i = 0, io = 0;
    while( uiQpos + (uint64_t)A.data[ i ].second < uiS )
    {
        if( A.data[ i ].first != MT_DELETION )
            uiQpos += (uint64_t)A.data[ i ].second;
        i++;
    }
    while( uiQposO + (uint64_t)O.data[ io ].second < uiS )
    {
        if( O.data[ io ].first != MT_DELETION )
            uiQposO += (uint64_t)O.data[ io ].second;
        io++;
    }
    while( uiQpos < uiE && uiQposO < uiE && i < A.data.size( ) && io < O.data.size( ) )
    {
        const uint64_t l = A.data[ i ].first != MT_DELETION ? (uint64_t)A.data[ i ].second : 0;
        const uint64_t lo = O.data[ io ].first != MT_DELETION ? (uint64_t)O.data[ io ].second : 0;
        const uint64_t s = std::max( std::max( uiQpos, uiQposO ), uiS );
        const uint64_t e = std::min( std::min( uiQpos + l, uiQposO + lo ), uiE );
        const uint64_t cur = s < e ? e - s : 0;
        if( A.data[ i ].first != MT_INSERTION && O.data[ io ].first != MT_INSERTION )
            uiOverlap += cur;
        if( uiQpos + l < uiQposO + lo )
            uiQpos += l, i++;
        else
            uiQposO += lo, io++;
    }
    const uint64_t uiSize = (uint64_t)std::min( A.end_q - A.begin_q, O.end_q - O.begin_q );
    return uiOverlap / static_cast<double>( uiSize );
}

std::vector<MqAln> mapping_quality( const Params& P, const std::vector<Alignment>& alns, int64_t qlen )
{
    std::vector<MqAln> v( alns.size( ) );
    for( size_t i = 0; i < alns.size( ); i++ )
        v[ i ].idx = (int)i;
    std::sort( v.begin( ), v.end( ), [ & ]( MqAln& a, MqAln& b ) { return alns[ a.idx ].score > alns[ b.idx ].score; } );
    if( v.empty( ) )
        return v;
    MqAln& first = v[ 0 ];
    const Alignment& F = alns[ first.idx ];
    first.secondary = false;
    size_t nSupp = 0;
    for( size_t i = 1; i < v.size( ); i++ )
    {
        v[ i ].mapq = 0.0;
        if( nSupp < (size_t)P.max_supplementary_per_prim &&
            overlap( alns[ v[ i ].idx ], F ) < P.max_overlap_supplementary )
            v[ i ].supplementary = true, v[ i ].secondary = false, nSupp++;
        else
            v[ i ].supplementary = false, v[ i ].secondary = true;
    }
    if( v.size( ) - nSupp >= 2 )
    {
        size_t k = 1;
        while( v[ k ].supplementary )
            k++;
        const int64_t s1 = F.score, s2 = alns[ v[ k ].idx ].score;
        if( s1 == 0 )
            first.mapq = 0;
        else
            first.mapq = static_cast<double>( s1 - s2 ) / static_cast<double>( s1 );
    }
    else
        first.mapq = F.score / (double)( (uint64_t)P.match * (uint64_t)qlen );
    if( num_seeds( F ) <= 1 )
        first.mapq /= 2;
    if( (double)F.score >= (double)( (uint64_t)P.match * (uint64_t)qlen ) * 0.8 && v.size( ) >= 3 )
        first.mapq *= 2;
    if( first.mapq > 1 )
        first.mapq = 1;
    if( nSupp > 0 )
    {
        for( size_t i = 1; i < v.size( ); i++ )
            if( v[ i ].supplementary )
                v[ i ].mapq = v[ 0 ].mapq;
        std::sort( v.begin( ), v.end( ), [ & ]( MqAln a, MqAln b ) { // Alignment::larger
            const size_t uiA = a.supplementary ? 1 : a.secondary ? 2 : 0, uiB = b.supplementary ? 1 : b.secondary ? 2 : 0;
            if( uiA != uiB )
                return uiA < uiB;
            const int64_t sa = alns[ a.idx ].score, sb = alns[ b.idx ].score;
            if( sa == sb )
                return alns[ a.idx ].soc_index < alns[ b.idx ].soc_index;
            return sa > sb;
        } );
    }
    if( P.report_n != 0 && v.size( ) > (size_t)P.report_n + nSupp )
        v.erase( v.begin( ) + P.report_n + nSupp, v.end( ) );
    v.erase( std::remove_if( v.begin( ), v.end( ),
                             [ & ]( const MqAln& a ) { return alns[ a.idx ].score < (long)P.min_alignment_score; } ),
             v.end( ) );
    return v;
}

std::vector<PairAln> paired_reads( const Index& I, const Params& P, const std::vector<Alignment>& alns1,
                                   std::vector<MqAln>& mq1, int64_t qlen1, const std::vector<Alignment>& alns2,
                                   std::vector<MqAln>& mq2, int64_t qlen2 )
{
    std::vector<PairAln> ret;
    if( mq1.empty( ) )
    {
        for( auto& m : mq2 )
            ret.push_back( PairAln{ 1, m } );
        return ret;
    }
    if( mq2.empty( ) )
    {
        for( auto& m : mq1 )
            ret.push_back( PairAln{ 0, m } );
        return ret;
    }
    std::vector<std::tuple<int64_t, bool, size_t, size_t>> vScores;
    const size_t mean = (size_t)P.paired_mean;
    for( size_t i = 0; i < mq1.size( ); i++ )
    {
        const Alignment& A1 = alns1[ mq1[ i ].idx ];
        if( A1.length == 0 )
            continue;
        for( size_t j = 0; j < mq2.size( ); j++ )
        {
            const Alignment& A2 = alns2[ mq2[ j ].idx ];
            if( A2.length == 0 )
                continue;
            int64_t iScore = A1.score + A2.score;
            bool bIsPaired = false;
            if( I.onReverse( A1.begin_ref ) != I.onReverse( A2.begin_ref ) )
            {
                const uint64_t uiP1 = (uint64_t)A1.begin_ref;
                const uint64_t uiP2 = (uint64_t)I.ref_len - ( (uint64_t)A2.begin_ref + 1 );
                const uint64_t d = uiP1 < uiP2 ? uiP2 - uiP1 : uiP1 - uiP2;
                if( ( (double)d ) >= ( (double)mean ) - P.paired_std * 3 &&
                    ( (double)d ) <= ( (double)mean ) + P.paired_std * 3 )
                {
                    iScore = ( int64_t )( iScore * P.paired_bonus );
                    bIsPaired = true;
                }
            }
            vScores.emplace_back( iScore, bIsPaired, i, j );
        }
    }
    std::sort( vScores.begin( ), vScores.end( ),
               []( const std::tuple<int64_t, bool, size_t, size_t>& rtA,
                   const std::tuple<int64_t, bool, size_t, size_t>& rtB ) {
                   if( std::get<0>( rtA ) == std::get<0>( rtB ) )
                       return std::get<1>( rtA ) && !std::get<1>( rtB );
                   return std::get<0>( rtA ) > std::get<0>( rtB );
               } );
    if( vScores.empty( ) )
        throw std::runtime_error( "paired_reads: no candidate pair (the reference reads vScores[0] here)" );
    const size_t i1 = std::get<2>( vScores[ 0 ] ), i2 = std::get<3>( vScores[ 0 ] );
    mq1[ i1 ].secondary = mq2[ i2 ].secondary = false;
    mq1[ i1 ].supplementary = mq2[ i2 ].supplementary = false;
    if( std::get<1>( vScores[ 0 ] ) && vScores.size( ) > 1 )
    {
        float fMapQ = ( (float)( std::get<0>( vScores[ 0 ] ) - std::get<0>( vScores[ 1 ] ) ) ) / std::get<0>( vScores[ 0 ] );
        const Alignment &A1 = alns1[ mq1[ i1 ].idx ], &A2 = alns2[ mq2[ i2 ].idx ];
        if( num_seeds( A1 ) <= 1 && num_seeds( A2 ) <= 1 )
            fMapQ /= 2;
        if( A1.score >= P.match * (uint64_t)qlen1 * 0.8 && mq1.size( ) >= 3 )
            fMapQ *= 2;
        else if( A2.score >= P.match * (uint64_t)qlen2 * 0.8 && mq2.size( ) >= 3 )
            fMapQ *= 2;
        if( fMapQ > 1 )
            fMapQ = 1;
        mq1[ i1 ].mapq = fMapQ;
        mq2[ i2 ].mapq = fMapQ;
    }
    ret.push_back( PairAln{ 0, mq1[ i1 ] } );
    ret.push_back( PairAln{ 1, mq2[ i2 ] } );
    return ret;
}

} // namespace oracle
