// NeedlemanWunsch module glue for ONE harmonized seed set, host+device: reference window, DP task enumeration,
// stitching of the DP results into the alignment run list, scoring, dangling-indel removal.
//
// Replaces, bit-exactly:
//   NeedlemanWunsch::execute_one / dynPrg / ksw / ksw_dual_ext / ksw_ext / ksw_simplified
//                                   libs/ma/src/module/needlemanWunsch.cpp:24-79, 82-169, 239-497, 499-622, 625-877
//   Alignment::append / removeDangeling / larger    libs/ma/src/container/alignment.cpp:11-98, 240-295,
//                                                   libs/ma/inc/ma/container/alignment.h:819-842
//   Pack::vExtract (virtual reverse strand)          libs/ma/inc/ma/container/pack.h:1147-1236, 1429-1448
//
// The same walk over the seed set is executed twice: once with a planning visitor that emits the DP problems of the
// set (all of them are known before any DP runs: gaps depend on the seeds only), once — after the batched DP kernel —
// with an assembling visitor that consumes the DP results in the same order.
#pragma once
#include "ksw_types.cuh"
#include "sochar.cuh"

namespace ma
{

struct NwParams
{
    int match, mismatch, gap, extend, sv_penalty;
    int max_gap_area, padding, bandwidth_ext, min_bandwidth_gap, zdrop;
    // 1: a 1 x 1 gap between two seeds needs no DP call. The reference's global kswcpp call on one query and one
    // reference base returns the CIGAR 1M whenever it computes at all: the only alternative, an insertion plus a
    // deletion, costs at least 2 (q + e), and a substitution costlier than that makes kswcpp return before the first
    // cell (kswcpp_core.h:411-412, "-min_sc > 2 (q + e)") - the one case in which this flag is 0.
    int one_by_one_is_match = 0;
};

enum
{
    MT_SEED = 0,
    MT_MATCH = 1,
    MT_MISSMATCH = 2,
    MT_INSERTION = 3,
    MT_DELETION = 4
};

struct DAln // 80 bytes
{
    long long begin_ref, end_ref, score;
    int begin_q, end_q;
    int length, n_runs;
    unsigned int soc_index;
    int read;
    long long run_off; // into the run slab; run word = len << 3 | MatchType
    int rank; // position among the read's alignments after the reference's final std::sort
    int flags; // MA_ALN_* (mapq.cuh), set by the MappingQuality / PairedReads stage
    double mapq; // Alignment::fMappingQuality
    int rank_mq; // position in MappingQuality's result, -1: not reported
    int pair_rank; // position in PairedReads' result, -1: not in it
};

struct NwWindow
{
    unsigned long long beginRef, endRef; // virtual text coordinates
    bool valid;
};

// needlemanWunsch.cpp:652-733 (bLocal == false)
MA_HD inline NwWindow nw_window( const DevIndex& I, const NwParams& P, const DSeed* S, int n )
{
    NwWindow w{ 0, 0, false };
    if( n == 0 )
        return w;
    unsigned long long beginRef = (unsigned long long)S[ 0 ].r, endRef = (unsigned long long)( S[ n - 1 ].r + S[ n - 1 ].len );
    for( int i = 0; i < n; i++ )
    {
        const unsigned long long er = (unsigned long long)( S[ i ].r + S[ i ].len );
        if( endRef < er )
            endRef = er;
        if( beginRef > (unsigned long long)S[ i ].r )
            beginRef = (unsigned long long)S[ i ].r;
    }
    if( beginRef >= endRef || bridging_subsection( I, (long long)beginRef, (long long)( endRef - beginRef + 1 ) ) )
        return w;
    const unsigned long long total = 2ull * (unsigned long long)I.fwd_len;
    const long long iOldContig = seq_id_or_rev( I, (long long)beginRef );
    beginRef -= (unsigned long long)P.padding;
    if( beginRef > endRef )
        beginRef = 0;
    endRef += (unsigned long long)P.padding;
    if( endRef >= total )
        endRef = total - 1;
    if( seq_id_or_rev( I, (long long)beginRef ) != iOldContig )
        beginRef = (unsigned long long)start_of_seq_or_rev( I, iOldContig );
    if( seq_id_or_rev( I, (long long)endRef ) != iOldContig )
        endRef = (unsigned long long)end_of_seq_or_rev( I, iOldContig ) - 1;
    w.beginRef = beginRef, w.endRef = endRef, w.valid = true;
    return w;
}

// Visitor interface:
//   void dyn( fromQ, toQ, fromR, toR, bLocalBeginning, bLocalEnd )   — dynPrg
//   void seed( len ), void del( n ), void ins( n )                   — Alignment::append of the respective type
template <class V> MA_HD inline void nw_walk( const DSeed* S, int n, int qlen, const NwWindow& w, V& v )
{
    typedef unsigned long long u64;
    const u64 beginRef = w.beginRef, endRef = w.endRef;
    v.dyn( 0, (u64)S[ 0 ].q, 0, (u64)S[ 0 ].r - beginRef, true, false );
    u64 endQ = (u64)( S[ 0 ].q + S[ 0 ].len ), endR = (u64)( S[ 0 ].r + S[ 0 ].len ) - beginRef;
    v.seed( (u64)S[ 0 ].len );
    for( int i = 1; i < n; i++ )
    {
        const DSeed& s = S[ i ];
        if( s.len == 0 )
            continue;
        u64 ovQ = endQ - (u64)s.q;
        if( (u64)s.q > endQ )
            ovQ = 0;
        u64 ovR = endR - ( (u64)s.r - beginRef );
        if( (u64)s.r > endR + beginRef )
            ovR = 0;
        const u64 len = (u64)s.len;
        const u64 overlap = ovQ > ovR ? ovQ : ovR;
        if( len > overlap )
        {
            v.dyn( endQ, (u64)s.q, endR, (u64)s.r - beginRef, false, false );
            if( ovQ > ovR )
                v.del( ovQ - ovR );
            if( ovR > ovQ )
                v.ins( ovR - ovQ );
            v.seed( len - overlap );
            if( (u64)( s.q + s.len ) > endQ )
                endQ = (u64)( s.q + s.len );
            if( (u64)( s.r + s.len ) > endR + beginRef )
                endR = (u64)( s.r + s.len ) - beginRef;
        }
    }
    v.dyn( endQ, (u64)qlen - 1, endR, endRef - beginRef - 1, false, true );
}

// ---- planning --------------------------------------------------------------------------------------------------
struct NwPlanner
{
    const NwParams& P;
    KswTask* tasks; // nullptr: count only
    int n = 0;
    long long qbase; // offset of the read's base 0 in the read slab
    long long rbase; // window.beginRef (virtual text)
    MA_HD NwPlanner( const NwParams& P, KswTask* tasks, long long qbase, long long rbase )
        : P( P ), tasks( tasks ), qbase( qbase ), rbase( rbase )
    {}
    MA_HD void emit( unsigned long long fq, unsigned long long tq, unsigned long long fr, unsigned long long tr,
                     bool rev, int w, int zdrop, int flag )
    {
        if( tasks )
        {
            KswTask t;
            t.qlen = (int)( tq - fq ), t.tlen = (int)( tr - fr );
            t.qoff = rev ? qbase + (long long)tq - 1 : qbase + (long long)fq;
            t.toff = rev ? rbase + (long long)tr - 1 : rbase + (long long)fr;
            t.w = w, t.zdrop = zdrop, t.flag = flag;
            t.tag = MA_TASK_TPACK | ( rev ? ( MA_TASK_QREV | MA_TASK_TREV ) : 0 );
            // extensions: the glue consumes max_q / max_t / CIGAR only (needlemanWunsch.cpp:262-267, 334-335, 575-579)
            if( flag & MA_KSW_EXTZ_ONLY )
                t.tag |= MA_TASK_EARLYSTOP;
            tasks[ n ] = t;
        }
        n++;
    }
    MA_HD void dyn( unsigned long long fq, unsigned long long tq, unsigned long long fr, unsigned long long tr,
                    bool bLocalBeginning, bool bLocalEnd )
    {
        if( tr <= fr || tq <= fq )
            return; // the three early returns of dynPrg need no DP
        if( !bLocalBeginning && !bLocalEnd )
        {
            if( tq - fq > (unsigned long long)P.max_gap_area || tr - fr > (unsigned long long)P.max_gap_area )
            {
                emit( fq, tq, fr, tr, false, P.bandwidth_ext, P.zdrop, MA_KSW_EXTZ_ONLY );
                emit( fq, tq, fr, tr, true, P.bandwidth_ext, P.zdrop,
                      MA_KSW_EXTZ_ONLY | MA_KSW_RIGHT | MA_KSW_REV_CIGAR );
            }
            else
            {
                const int qlen = (int)( tq - fq ), tlen = (int)( tr - fr );
                if( qlen == 1 && tlen == 1 && P.one_by_one_is_match )
                    return; // 64 % of the gap fills of Illumina reads: a substitution between two seeds
                int w = P.min_bandwidth_gap;
                const int d = tlen - qlen < 0 ? qlen - tlen : tlen - qlen;
                if( d + 10 > w )
                    w = d + 10;
                emit( fq, tq, fr, tr, false, w, -1, 0 );
            }
            return;
        }
        if( bLocalBeginning )
            emit( fq, tq, fr, tr, true, P.bandwidth_ext, P.zdrop, MA_KSW_EXTZ_ONLY | MA_KSW_RIGHT | MA_KSW_REV_CIGAR );
        else
            emit( fq, tq, fr, tr, false, P.bandwidth_ext, P.zdrop, MA_KSW_EXTZ_ONLY );
    }
    MA_HD void seed( unsigned long long )
    {}
    MA_HD void del( unsigned long long )
    {}
    MA_HD void ins( unsigned long long )
    {}
};

// ---- assembling ------------------------------------------------------------------------------------------------
struct NwAssembler
{
    typedef unsigned long long u64;
    const DevIndex& I;
    const NwParams& P;
    const unsigned char* q; // the read, base 0
    const u64 beginRef;
    const KswOut* res; // DP results of this set, in planning order
    const unsigned int* cigar; // cigar slab
    int next = 0;
    // alignment under construction
    unsigned int* runs;
    int cap, nRuns = 0, front = 0;
    bool overflow = false;
    long long score = 0;
    u64 beginR, endR, beginQ = 0, endQ = 0, length = 0;

    MA_HD NwAssembler( const DevIndex& I, const NwParams& P, const unsigned char* q, u64 beginRef, const KswOut* res,
                       const unsigned int* cigar, unsigned int* runs, int cap )
        : I( I ), P( P ), q( q ), beginRef( beginRef ), res( res ), cigar( cigar ), runs( runs ), cap( cap ),
          beginR( beginRef ), endR( beginRef )
    {}
    MA_HD int refBase( u64 pos ) const // (*pRef)[pos] of the extracted window
    {
        return pack_virtual( I, (long long)( beginRef + pos ) );
    }
    MA_HD u64 gapPenalty( u64 n ) const
    {
        const u64 p = (u64)P.extend * n + (u64)P.gap;
        return p < (u64)P.sv_penalty ? p : (u64)P.sv_penalty;
    }
    // Alignment::append (alignment.cpp:11-98)
    MA_HD void append( int type, u64 size )
    {
        if( size == 0 )
            return;
        if( type == MT_SEED || type == MT_MATCH )
        {
            score += (long long)( (u64)P.match * size );
            endR += size, endQ += size;
        }
        else if( type == MT_MISSMATCH )
        {
            score -= (long long)( (u64)P.mismatch * size );
            endR += size, endQ += size;
        }
        else
        {
            if( type == MT_INSERTION )
                endQ += size;
            else
                endR += size;
            if( nRuns > front && (int)( runs[ nRuns - 1 ] & 7 ) == type )
            {
                const u64 last = runs[ nRuns - 1 ] >> 3;
                size += last;
                length -= last;
                score += (long long)gapPenalty( last );
                nRuns--;
            }
            score -= (long long)gapPenalty( size );
        }
        if( nRuns > front && (int)( runs[ nRuns - 1 ] & 7 ) == type )
            runs[ nRuns - 1 ] += (unsigned int)( size << 3 );
        else
        {
            if( nRuns < cap )
                runs[ nRuns ] = (unsigned int)( size << 3 ) | (unsigned int)type;
            else
                overflow = true;
            if( nRuns < cap )
                nRuns++;
        }
        length += size;
    }
    // the M run of a CIGAR split into match / missmatch runs (needlemanWunsch.cpp:131-144 appends base by base; k
    // appends of one type are the same integer additions as one append of k, so equal neighbours are appended together)
    MA_HD void matches( u64 qPos, u64 rPos, unsigned int n )
    {
        unsigned int runLen = 0;
        int runType = MT_MATCH;
        for( unsigned int i = 0; i < n; i++ )
        {
            const int t = q[ qPos + i ] == refBase( rPos + i ) ? MT_MATCH : MT_MISSMATCH;
            if( t != runType && runLen )
                append( runType, runLen ), runLen = 0;
            runType = t, runLen++;
        }
        append( runType, runLen );
    }
    MA_HD void seed( u64 n )
    {
        append( MT_SEED, n );
    }
    MA_HD void del( u64 n )
    {
        append( MT_DELETION, n );
    }
    MA_HD void ins( u64 n )
    {
        append( MT_INSERTION, n );
    }
    MA_HD void plain( const KswOut& ez, u64& qPos, u64& rPos ) // the cigar read-out loops of ksw() / dynPrg()
    {
        const unsigned int* c = cigar + ez.cigar_off;
        for( int i = 0; i < ez.n_cigar; ++i )
        {
            const unsigned int sym = c[ i ] & 0xf, amount = c[ i ] >> 4;
            if( sym == 0 )
            {
                matches( qPos, rPos, amount );
                qPos += amount, rPos += amount;
            }
            else if( sym == 1 )
                append( MT_INSERTION, amount ), qPos += amount;
            else
                append( MT_DELETION, amount ), rPos += amount;
        }
    }
    // needlemanWunsch.cpp:239-497
    MA_HD void dual( u64 fromQuery, u64 toQuery, u64 fromRef, u64 toRef )
    {
        const KswOut& L = res[ next++ ];
        const KswOut& R = res[ next++ ];
        const unsigned int* cl = cigar + L.cigar_off;
        const unsigned int* cr = cigar + R.cigar_off;
        u64 qCenter = ( fromQuery + (u64)(long long)L.max_q + ( toQuery - (u64)(long long)R.max_q - 1 ) ) / 2;
        {
            const u64 m = toQuery < qCenter ? toQuery : qCenter;
            qCenter = fromQuery > m ? fromQuery : m;
        }
        u64 rCenter = ( fromRef + (u64)(long long)L.max_t + ( toRef - (u64)(long long)R.max_t - 1 ) ) / 2;
        {
            const u64 m = toRef < rCenter ? toRef : rCenter;
            rCenter = fromRef > m ? fromRef : m;
        }
        u64 qPos = fromQuery, rPos = fromRef;
        if( rPos != rCenter && qPos != qCenter )
            for( int i = 0; i < L.n_cigar; ++i )
            {
                const unsigned int sym = cl[ i ] & 0xf;
                unsigned int amount = cl[ i ] >> 4;
                if( sym == 0 )
                {
                    if( qPos + amount > qCenter )
                        amount = (unsigned int)( qCenter - qPos );
                    if( rPos + amount > rCenter )
                        amount = (unsigned int)( rCenter - rPos );
                    matches( qPos, rPos, amount );
                    qPos += amount, rPos += amount;
                }
                else if( sym == 1 )
                {
                    if( qPos + amount > qCenter )
                        amount = (unsigned int)( qCenter - qPos );
                    append( MT_INSERTION, amount );
                    qPos += amount;
                }
                else
                {
                    if( rPos + amount > rCenter )
                        amount = (unsigned int)( rCenter - rPos );
                    append( MT_DELETION, amount );
                    rPos += amount;
                }
                if( rPos == rCenter )
                    break;
                if( qPos == qCenter )
                    break;
            }
        u64 rPosRight = toRef - (u64)(long long)R.max_t - 1, qPosRight = toQuery - (u64)(long long)R.max_q - 1;
        unsigned int notUnrolled = 0;
        int lastType = MT_SEED;
        int i = 0;
        for( ; i < R.n_cigar; ++i )
        {
            if( rPosRight >= rCenter && qPosRight >= qCenter )
                break;
            const unsigned int sym = cr[ i ] & 0xf;
            unsigned int amount = cr[ i ] >> 4;
            if( sym == 0 )
            {
                if( rPosRight + amount >= rCenter && qPosRight + amount >= qCenter )
                {
                    if( rPosRight < rCenter && ( qPosRight >= qCenter || rCenter - rPosRight > qCenter - qPosRight ) )
                    {
                        notUnrolled = amount - (unsigned int)( rCenter - rPosRight );
                        amount = (unsigned int)( rCenter - rPosRight );
                    }
                    else
                    {
                        notUnrolled = amount - (unsigned int)( qCenter - qPosRight );
                        amount = (unsigned int)( qCenter - qPosRight );
                    }
                }
                qPosRight += amount, rPosRight += amount;
                lastType = MT_MATCH;
            }
            else if( sym == 1 )
            {
                if( qPosRight + amount > qCenter && rPosRight >= rCenter )
                {
                    notUnrolled = amount - (unsigned int)( qCenter - qPosRight );
                    amount = (unsigned int)( qCenter - qPosRight );
                }
                qPosRight += amount;
                lastType = MT_INSERTION;
            }
            else
            {
                if( rPosRight + amount > rCenter && qPosRight >= qCenter )
                {
                    notUnrolled = amount - (unsigned int)( rCenter - rPosRight );
                    amount = (unsigned int)( rCenter - rPosRight );
                }
                rPosRight += amount;
                lastType = MT_DELETION;
            }
        }
        // :404-432 — unsigned arithmetic and the operator-precedence quirk are kept
        const u64 dq = qPosRight - qPos, dr = rPosRight - rPos;
        u64 uiMMPenalty = dq >= dr ? dq - dr : dr - dq;
        uiMMPenalty *= (u64)P.mismatch;
        const u64 uiM = dq < dr ? dq : dr;
        const long long kq = (signed char)P.gap, ke = (signed char)P.extend;
        if( uiM > 0 )
            uiMMPenalty += (u64)kq + (u64)ke * uiM;
        u64 uiGapPenalty = 0;
        if( dq > 0 )
            uiGapPenalty += (u64)kq + (u64)ke * qPosRight - qPos;
        if( dr > 0 )
            uiGapPenalty += (u64)kq + (u64)ke * rPosRight - rPos;
        if( uiMMPenalty < uiGapPenalty )
            while( qPos < qPosRight && rPos < rPosRight )
            {
                append( q[ qPos ] == refBase( rPos ) ? MT_MATCH : MT_MISSMATCH, 1 );
                qPos++, rPos++;
            }
        append( MT_INSERTION, qPosRight - qPos );
        append( MT_DELETION, rPosRight - rPos );
        if( lastType == MT_MATCH )
            matches( qPosRight, rPosRight, notUnrolled );
        else
            append( lastType, notUnrolled );
        if( lastType == MT_MATCH )
            qPosRight += notUnrolled, rPosRight += notUnrolled;
        else if( lastType == MT_INSERTION )
            qPosRight += notUnrolled;
        else if( lastType == MT_DELETION )
            rPosRight += notUnrolled;
        for( ; i < R.n_cigar; ++i )
        {
            const unsigned int sym = cr[ i ] & 0xf, amount = cr[ i ] >> 4;
            if( sym == 0 )
            {
                matches( qPosRight, rPosRight, amount );
                qPosRight += amount, rPosRight += amount;
            }
            else if( sym == 1 )
                append( MT_INSERTION, amount ), qPosRight += amount;
            else
                append( MT_DELETION, amount ), rPosRight += amount;
        }
    }
    // dynPrg (needlemanWunsch.cpp:499-622) incl. ksw (:82-169)
    MA_HD void dyn( u64 fromQuery, u64 toQuery, u64 fromRef, u64 toRef, bool bLocalBeginning, bool bLocalEnd )
    {
        if( toRef <= fromRef )
            if( toQuery <= fromQuery )
                return;
        if( toQuery <= fromQuery )
        {
            append( MT_DELETION, toRef - fromRef );
            return;
        }
        if( toRef <= fromRef )
        {
            append( MT_INSERTION, toQuery - fromQuery );
            return;
        }
        if( !bLocalBeginning && !bLocalEnd )
        {
            if( toQuery - fromQuery > (u64)P.max_gap_area || toRef - fromRef > (u64)P.max_gap_area )
                dual( fromQuery, toQuery, fromRef, toRef );
            else if( toQuery - fromQuery == 1 && toRef - fromRef == 1 && P.one_by_one_is_match )
                matches( fromQuery, fromRef, 1 ); // the CIGAR 1M of the reference's call, no leftovers
            else
            {
                const KswOut& ez = res[ next++ ];
                u64 qPos = fromQuery, rPos = fromRef;
                plain( ez, qPos, rPos );
                // sic: leftovers are appended with swapped types (:167-168)
                append( MT_DELETION, toQuery - qPos );
                append( MT_INSERTION, toRef - rPos );
            }
            return;
        }
        const bool bReverse = bLocalBeginning;
        const KswOut& ez = res[ next++ ];
        u64 qPos = fromQuery, rPos = fromRef;
        if( bReverse )
        {
            rPos = toRef - (u64)(long long)ez.max_t - 1;
            qPos = toQuery - (u64)(long long)ez.max_q - 1;
        }
        plain( ez, qPos, rPos );
        if( bReverse )
        {
            const u64 sr = toRef - (u64)(long long)ez.max_t - 1, sq = toQuery - (u64)(long long)ez.max_q - 1;
            beginR += sr, endR += sr;
            beginQ += sq, endQ += sq;
        }
    }
    // Alignment::removeDangeling (alignment.cpp:240-295)
    MA_HD void removeDangeling( )
    {
        if( nRuns == front )
            return;
        while( nRuns > front && ( ( runs[ front ] & 7 ) == MT_DELETION || ( runs[ front ] & 7 ) == MT_INSERTION ) )
        {
            const u64 n = runs[ front ] >> 3;
            if( ( runs[ front ] & 7 ) == MT_DELETION )
                beginR += n;
            else
                beginQ += n;
            score += (long long)gapPenalty( n );
            length -= n;
            front++;
        }
        while( nRuns > front &&
               ( ( runs[ nRuns - 1 ] & 7 ) == MT_DELETION || ( runs[ nRuns - 1 ] & 7 ) == MT_INSERTION ) )
        {
            const u64 n = runs[ nRuns - 1 ] >> 3;
            if( ( runs[ nRuns - 1 ] & 7 ) == MT_DELETION )
                endR -= n;
            else
                endQ -= n;
            score += (long long)gapPenalty( n );
            length -= n;
            nRuns--;
        }
    }
};

} // namespace ma
