// CPU ORACLE — test infrastructure only (see oracle.h).
// NeedlemanWunsch module glue restated from
//   libs/ma/inc/ma/module/needlemanWunsch.h:111-134 (execute), libs/ma/src/module/needlemanWunsch.cpp:24-79 (ksw_ext,
//   ksw_simplified), :82-169 (ksw), :239-497 (ksw_dual_ext), :499-622 (dynPrg), :625-877 (execute_one)
//   libs/ma/src/container/alignment.cpp:11-98 (append), :240-295 (removeDangeling), alignment.h:819-842 (larger)
#include "ma_oracle.h"
#include <algorithm>
#include <stdexcept>

namespace oracle
{
namespace
{
struct NW
{
    const Index& I;
    const Params& P;
    std::vector<KswCall>* pLog;
    ma_oracle_score_t sc;
    NW( const Index& I, const Params& P, std::vector<KswCall>* pLog ) : I( I ), P( P ), pLog( pLog )
    {
        sc = ma_oracle_score_t{ P.match, P.mismatch, P.gap, P.extend, P.gap2, P.extend2 };
    }

    // alignment.cpp:11-98
    void append( Alignment& A, int type, int64_t size )
    {
        if( size == 0 )
            return;
        uint64_t usize = (uint64_t)size;
        if( type == MT_SEED || type == MT_MATCH )
        {
            A.score += P.match * usize;
            A.end_ref += usize, A.end_q += usize;
        }
        else if( type == MT_MISSMATCH )
        {
            A.score -= P.mismatch * usize;
            A.end_ref += usize, A.end_q += usize;
        }
        else
        {
            if( type == MT_INSERTION )
                A.end_q += usize;
            else
                A.end_ref += usize;
            if( !A.data.empty( ) && A.data.back( ).first == type )
            {
                usize += A.data.back( ).second;
                A.length -= A.data.back( ).second;
                if( (uint64_t)( P.extend * A.data.back( ).second + P.gap ) < (uint64_t)P.sv_penalty )
                    A.score += P.extend * A.data.back( ).second + P.gap;
                else
                    A.score += (uint64_t)P.sv_penalty;
                A.data.pop_back( );
            }
            if( (uint64_t)( P.extend * usize + P.gap ) < (uint64_t)P.sv_penalty )
                A.score -= P.extend * usize + P.gap;
            else
                A.score -= (uint64_t)P.sv_penalty;
        }
        if( !A.data.empty( ) && A.data.back( ).first == type )
            A.data.back( ).second += usize;
        else
            A.data.push_back( { type, (int64_t)usize } );
        A.length += usize;
    }

    void appendMatches( Alignment& A, const std::vector<uint8_t>& q, const std::vector<uint8_t>& r, uint64_t qPos,
                        uint64_t rPos, uint32_t n )
    {
        for( uint32_t i = 0; i < n; i++ )
            append( A, q[ qPos + i ] == r[ rPos + i ] ? MT_MATCH : MT_MISSMATCH, 1 );
    }

    struct Ez
    {
        ma_oracle_ksw_t ez;
        std::vector<uint32_t> cigar;
    };

    Ez runKsw( const uint8_t* q, int qlen, const uint8_t* t, int tlen, int w, int zdrop, int flag )
    {
        Ez r;
        r.cigar.resize( (size_t)qlen + tlen + 8 );
        int64_t cells = 0;
        if( ma_oracle_ksw( qlen, q, tlen, t, &sc, w, zdrop, flag, &r.ez, r.cigar.data( ), (int)r.cigar.size( ),
                           &cells ) )
            throw std::runtime_error( "oracle ksw failed" );
        r.cigar.resize( r.ez.n_cigar );
        if( pLog )
        {
            KswCall c;
            int64_t a[ 16 ] = { qlen,       tlen,        w,        zdrop,       flag,      r.ez.max,
                                r.ez.zdropped, r.ez.max_q, r.ez.max_t, r.ez.mqe,  r.ez.mqe_t, r.ez.mte,
                                r.ez.mte_q, r.ez.score,  r.ez.n_cigar, r.ez.reach_end };
            std::copy( a, a + 16, c.f );
            c.q.assign( q, q + qlen );
            c.t.assign( t, t + tlen );
            c.cigar = r.cigar;
            pLog->push_back( c );
        }
        return r;
    }

    Ez ksw_ext( std::vector<uint8_t>& q, uint64_t fq, uint64_t tq, std::vector<uint8_t>& r, uint64_t fr, uint64_t tr,
                bool bRef ) // needlemanWunsch.cpp:24-55
    {
        return runKsw( q.data( ) + fq, (int)( tq - fq ), r.data( ) + fr, (int)( tr - fr ), P.bandwidth_ext, P.zdrop,
                       bRef ? ( MA_KSW_EXTZ_ONLY | MA_KSW_RIGHT | MA_KSW_REV_CIGAR ) : MA_KSW_EXTZ_ONLY );
    }

    // needlemanWunsch.cpp:82-169
    void ksw( std::vector<uint8_t>& q, std::vector<uint8_t>& r, uint64_t fq, uint64_t tq, uint64_t fr, uint64_t tr,
              Alignment& A )
    {
        if( tr <= fr )
            if( tq <= fq )
                return;
        if( tq <= fq )
        {
            append( A, MT_DELETION, tr - fr );
            return;
        }
        if( tr <= fr )
        {
            append( A, MT_INSERTION, tq - fq );
            return;
        }
        int qlen = (int)( tq - fq ), tlen = (int)( tr - fr ), w = P.min_bandwidth_gap;
        if( std::abs( tlen - qlen ) + 10 > w ) // ksw_simplified, :57-79
            w = std::abs( tlen - qlen ) + 10;
        Ez ez = runKsw( q.data( ) + fq, qlen, r.data( ) + fr, tlen, w, -1, 0 );
        uint64_t qPos = fq, rPos = fr;
        for( uint32_t c : ez.cigar )
        {
            uint32_t sym = c & 0xf, amount = c >> 4;
            switch( sym )
            {
                case 0:
                    appendMatches( A, q, r, qPos, rPos, amount );
                    qPos += amount, rPos += amount;
                    break;
                case 1: append( A, MT_INSERTION, amount ), qPos += amount; break;
                case 2: append( A, MT_DELETION, amount ), rPos += amount; break;
            }
        }
        // sic: the leftovers are appended with swapped types (:167-168)
        append( A, MT_DELETION, tq - qPos );
        append( A, MT_INSERTION, tr - rPos );
    }

    // needlemanWunsch.cpp:239-497
    void ksw_dual_ext( std::vector<uint8_t>& q, std::vector<uint8_t>& r, uint64_t fromQuery, uint64_t toQuery,
                       uint64_t fromRef, uint64_t toRef, Alignment& A )
    {
        Ez L = ksw_ext( q, fromQuery, toQuery, r, fromRef, toRef, false );
        std::reverse( q.begin( ) + fromQuery, q.begin( ) + toQuery );
        std::reverse( r.begin( ) + fromRef, r.begin( ) + toRef );
        Ez R = ksw_ext( q, fromQuery, toQuery, r, fromRef, toRef, true );
        std::reverse( q.begin( ) + fromQuery, q.begin( ) + toQuery );
        std::reverse( r.begin( ) + fromRef, r.begin( ) + toRef );

        uint64_t qCenter = ( fromQuery + L.ez.max_q + ( toQuery - R.ez.max_q - 1 ) ) / 2;
        qCenter = std::max( fromQuery, std::min( toQuery, qCenter ) );
        uint64_t rCenter = ( fromRef + L.ez.max_t + ( toRef - R.ez.max_t - 1 ) ) / 2;
        rCenter = std::max( fromRef, std::min( toRef, rCenter ) );
        uint64_t qPos = fromQuery, rPos = fromRef;
        if( rPos != rCenter && qPos != qCenter )
            for( uint32_t c : L.cigar )
            {
                uint32_t sym = c & 0xf, amount = c >> 4;
                switch( sym )
                {
                    case 0:
                        if( qPos + amount > qCenter )
                            amount = (uint32_t)( qCenter - qPos );
                        if( rPos + amount > rCenter )
                            amount = (uint32_t)( rCenter - rPos );
                        appendMatches( A, q, r, qPos, rPos, amount );
                        qPos += amount, rPos += amount;
                        break;
                    case 1:
                        if( qPos + amount > qCenter )
                            amount = (uint32_t)( qCenter - qPos );
                        append( A, MT_INSERTION, amount );
                        qPos += amount;
                        break;
                    case 2:
                        if( rPos + amount > rCenter )
                            amount = (uint32_t)( rCenter - rPos );
                        append( A, MT_DELETION, amount );
                        rPos += amount;
                        break;
                }
                if( rPos == rCenter )
                    break;
                if( qPos == qCenter )
                    break;
            }
        uint64_t rPosRight = toRef - R.ez.max_t - 1, qPosRight = toQuery - R.ez.max_q - 1;
        uint32_t notUnrolled = 0;
        int lastType = MT_SEED;
        size_t i = 0;
        for( ; i < R.cigar.size( ); ++i )
        {
            if( rPosRight >= rCenter && qPosRight >= qCenter )
                break;
            uint32_t sym = R.cigar[ i ] & 0xf, amount = R.cigar[ i ] >> 4;
            switch( sym )
            {
                case 0:
                    if( rPosRight + amount >= rCenter && qPosRight + amount >= qCenter )
                    {
                        if( rPosRight < rCenter &&
                            ( qPosRight >= qCenter || rCenter - rPosRight > qCenter - qPosRight ) )
                        {
                            notUnrolled = amount - (uint32_t)( rCenter - rPosRight );
                            amount = (uint32_t)( rCenter - rPosRight );
                        }
                        else
                        {
                            notUnrolled = amount - (uint32_t)( qCenter - qPosRight );
                            amount = (uint32_t)( qCenter - qPosRight );
                        }
                    }
                    qPosRight += amount, rPosRight += amount;
                    lastType = MT_MATCH;
                    break;
                case 1:
                    if( qPosRight + amount > qCenter && rPosRight >= rCenter )
                    {
                        notUnrolled = amount - (uint32_t)( qCenter - qPosRight );
                        amount = (uint32_t)( qCenter - qPosRight );
                    }
                    qPosRight += amount;
                    lastType = MT_INSERTION;
                    break;
                case 2:
                    if( rPosRight + amount > rCenter && qPosRight >= qCenter )
                    {
                        notUnrolled = amount - (uint32_t)( rCenter - rPosRight );
                        amount = (uint32_t)( rCenter - rPosRight );
                    }
                    rPosRight += amount;
                    lastType = MT_DELETION;
                    break;
            }
        }
        // gap between the two extensions (:404-432), unsigned arithmetic and the precedence quirk kept
        const uint64_t uiMissMatch = (uint64_t)P.mismatch;
        const int8_t kq = (int8_t)P.gap, ke = (int8_t)P.extend;
        uint64_t uiMMPenalty = ( qPosRight - qPos ) >= ( rPosRight - rPos )
                                   ? ( qPosRight - qPos ) - ( rPosRight - rPos )
                                   : ( rPosRight - rPos ) - ( qPosRight - qPos );
        uiMMPenalty *= uiMissMatch;
        size_t uiM = std::min( ( qPosRight - qPos ), ( rPosRight - rPos ) );
        if( uiM > 0 )
            uiMMPenalty += kq + ke * uiM;
        uint64_t uiGapPenalty = 0;
        if( qPosRight - qPos > 0 )
            uiGapPenalty += kq + ke * qPosRight - qPos;
        if( rPosRight - rPos > 0 )
            uiGapPenalty += kq + ke * rPosRight - rPos;
        if( uiMMPenalty < uiGapPenalty )
            while( qPos < qPosRight && rPos < rPosRight )
            {
                append( A, q[ qPos ] == r[ rPos ] ? MT_MATCH : MT_MISSMATCH, 1 );
                qPos++, rPos++;
            }
        append( A, MT_INSERTION, qPosRight - qPos );
        append( A, MT_DELETION, rPosRight - rPos );
        if( lastType == MT_MATCH )
            appendMatches( A, q, r, qPosRight, rPosRight, notUnrolled );
        else
            append( A, lastType, notUnrolled );
        switch( lastType )
        {
            case MT_MATCH: qPosRight += notUnrolled, rPosRight += notUnrolled; break;
            case MT_INSERTION: qPosRight += notUnrolled; break;
            case MT_DELETION: rPosRight += notUnrolled; break;
            default: break;
        }
        for( ; i < R.cigar.size( ); ++i )
        {
            uint32_t sym = R.cigar[ i ] & 0xf, amount = R.cigar[ i ] >> 4;
            switch( sym )
            {
                case 0:
                    appendMatches( A, q, r, qPosRight, rPosRight, amount );
                    qPosRight += amount, rPosRight += amount;
                    break;
                case 1: append( A, MT_INSERTION, amount ), qPosRight += amount; break;
                case 2: append( A, MT_DELETION, amount ), rPosRight += amount; break;
            }
        }
    }

    // needlemanWunsch.cpp:499-622
    void dynPrg( std::vector<uint8_t>& q, std::vector<uint8_t>& r, uint64_t fromQuery, uint64_t toQuery,
                 uint64_t fromRef, uint64_t toRef, Alignment& A, bool bLocalBeginning, bool bLocalEnd )
    {
        if( toRef <= fromRef )
            if( toQuery <= fromQuery )
                return;
        if( toQuery <= fromQuery )
        {
            append( A, MT_DELETION, toRef - fromRef );
            return;
        }
        if( toRef <= fromRef )
        {
            append( A, MT_INSERTION, toQuery - fromQuery );
            return;
        }
        if( !bLocalBeginning && !bLocalEnd )
        {
            if( toQuery - fromQuery > (uint64_t)P.max_gap_area || toRef - fromRef > (uint64_t)P.max_gap_area )
                ksw_dual_ext( q, r, fromQuery, toQuery, fromRef, toRef, A );
            else
                ksw( q, r, fromQuery, toQuery, fromRef, toRef, A );
            return;
        }
        const bool bReverse = bLocalBeginning;
        if( bReverse )
        {
            std::reverse( q.begin( ) + fromQuery, q.begin( ) + toQuery );
            std::reverse( r.begin( ) + fromRef, r.begin( ) + toRef );
        }
        Ez ez = ksw_ext( q, fromQuery, toQuery, r, fromRef, toRef, bReverse );
        if( bReverse )
        {
            std::reverse( q.begin( ) + fromQuery, q.begin( ) + toQuery );
            std::reverse( r.begin( ) + fromRef, r.begin( ) + toRef );
        }
        uint64_t qPos = fromQuery, rPos = fromRef;
        if( bReverse )
        {
            rPos = toRef - ez.ez.max_t - 1;
            qPos = toQuery - ez.ez.max_q - 1;
        }
        for( uint32_t c : ez.cigar )
        {
            uint32_t sym = c & 0xf, amount = c >> 4;
            switch( sym )
            {
                case 0:
                    appendMatches( A, q, r, qPos, rPos, amount );
                    qPos += amount, rPos += amount;
                    break;
                case 1: append( A, MT_INSERTION, amount ), qPos += amount; break;
                case 2: append( A, MT_DELETION, amount ), rPos += amount; break;
            }
        }
        if( bReverse )
        {
            const uint64_t sr = toRef - ez.ez.max_t - 1, sq = toQuery - ez.ez.max_q - 1;
            A.begin_ref += sr, A.end_ref += sr;
            A.begin_q += sq, A.end_q += sq;
        }
    }

    void removeDangeling( Alignment& A ) // alignment.cpp:240-295
    {
        if( A.data.empty( ) )
            return;
        auto pen = [ & ]( int64_t n ) -> uint64_t {
            if( (uint64_t)( P.gap + P.extend * (uint64_t)n ) < (uint64_t)P.sv_penalty )
                return P.gap + P.extend * (uint64_t)n;
            return (uint64_t)P.sv_penalty;
        };
        while( A.data.front( ).first == MT_DELETION || A.data.front( ).first == MT_INSERTION )
        {
            if( A.data.front( ).first == MT_DELETION )
                A.begin_ref += A.data.front( ).second;
            else
                A.begin_q += A.data.front( ).second;
            A.score += pen( A.data.front( ).second );
            A.length -= A.data.front( ).second;
            A.data.erase( A.data.begin( ) );
        }
        while( A.data.back( ).first == MT_DELETION || A.data.back( ).first == MT_INSERTION )
        {
            if( A.data.back( ).first == MT_DELETION )
                A.end_ref -= A.data.back( ).second;
            else
                A.end_q -= A.data.back( ).second;
            A.score += pen( A.data.back( ).second );
            A.length -= A.data.back( ).second;
            A.data.pop_back( );
        }
    }

    // needlemanWunsch.cpp:625-877 (bLocal = false)
    Alignment execute_one( SeedSet& set, std::vector<uint8_t>& query )
    {
        auto& S = set.seeds;
        Alignment A;
        A.soc_index = set.soc_index;
        if( S.empty( ) )
            return A;
        const uint64_t qlen = query.size( );
        uint64_t beginRef = S.front( ).r, endRef = S.back( ).end_ref( );
        uint64_t endQuery = S.back( ).end( ), beginQuery = S.front( ).q;
        for( auto& x : S )
        {
            if( endRef < (uint64_t)x.end_ref( ) )
                endRef = x.end_ref( );
            if( beginRef > (uint64_t)x.r )
                beginRef = x.r;
            if( endQuery < (uint64_t)x.end( ) )
                endQuery = x.end( );
            if( beginQuery > (uint64_t)x.end( ) )
                beginQuery = x.q;
        }
        if( beginRef >= endRef || I.bridging( beginRef, endRef - beginRef + 1 ) )
            return A;
        const uint64_t total = 2 * (uint64_t)I.fwd_len;
        int64_t iOldContig = I.seqIdForPositionOrRev( beginRef );
        beginRef -= (uint64_t)P.padding;
        if( beginRef > endRef )
            beginRef = 0;
        endRef += (uint64_t)P.padding;
        if( endRef >= total )
            endRef = total - 1;
        endQuery = qlen;
        beginQuery = 0;
        if( I.seqIdForPositionOrRev( beginRef ) != iOldContig )
            beginRef = I.startOfSeqOrRev( iOldContig );
        if( I.seqIdForPositionOrRev( endRef ) != iOldContig )
            endRef = I.endOfSeqOrRev( iOldContig ) - 1;
        A.begin_ref = A.end_ref = beginRef;
        A.begin_q = A.end_q = beginQuery;
        std::vector<uint8_t> ref;
        I.extract( beginRef, endRef, ref );
        dynPrg( query, ref, 0, S.front( ).q, 0, S.front( ).r - beginRef, A, true, false );
        uint64_t endOfLastSeedQuery = S.front( ).end( );
        uint64_t endOfLastSeedReference = S.front( ).end_ref( ) - beginRef;
        append( A, MT_SEED, S.front( ).len );
        bool bSkip = true;
        for( Seed& rSeed : S )
        {
            if( bSkip )
            {
                bSkip = false;
                continue;
            }
            if( rSeed.len == 0 )
                continue;
            uint64_t ovQ = endOfLastSeedQuery - rSeed.q;
            if( (uint64_t)rSeed.q > endOfLastSeedQuery )
                ovQ = 0;
            uint64_t ovR = endOfLastSeedReference - ( rSeed.r - beginRef );
            if( (uint64_t)rSeed.r > endOfLastSeedReference + beginRef )
                ovR = 0;
            uint64_t len = rSeed.len;
            uint64_t overlap = std::max( ovQ, ovR );
            if( len > overlap )
            {
                dynPrg( query, ref, endOfLastSeedQuery, rSeed.q, endOfLastSeedReference, rSeed.r - beginRef, A, false,
                        false );
                if( ovQ > ovR )
                    append( A, MT_DELETION, ovQ - ovR );
                if( ovR > ovQ )
                    append( A, MT_INSERTION, ovR - ovQ );
                append( A, MT_SEED, len - overlap );
                if( (uint64_t)rSeed.end( ) > endOfLastSeedQuery )
                    endOfLastSeedQuery = rSeed.end( );
                if( (uint64_t)rSeed.end_ref( ) > endOfLastSeedReference + beginRef )
                    endOfLastSeedReference = rSeed.end_ref( ) - beginRef;
            }
        }
        dynPrg( query, ref, endOfLastSeedQuery, endQuery - 1, endOfLastSeedReference, endRef - beginRef - 1, A, false,
                true );
        removeDangeling( A );
        return A;
    }
};
} // namespace

std::vector<Alignment> needleman_wunsch( const Index& I, const Params& P, std::vector<SeedSet>& sets,
                                         std::vector<uint8_t>& query, std::vector<KswCall>* pLog )
{
    NW nw( I, P, pLog );
    std::vector<Alignment> out;
    for( auto& s : sets )
        out.push_back( nw.execute_one( s, query ) );
    // needlemanWunsch.h:131-132 with Alignment::larger (bSecondary/bSupplementary are still false here)
    std::sort( out.begin( ), out.end( ), []( const Alignment& a, const Alignment& b ) {
        if( a.score == b.score )
            return a.soc_index < b.soc_index;
        return a.score > b.score;
    } );
    return out;
}

} // namespace oracle
