// FM-index / pack accessors and the per-read seeding routine (BinarySeeding + seed enumeration), host+device.
//
// Replaces, bit-exactly:
//   FMIndex::bwt_occ4 / bwt_2occ4      libs/ma/inc/ma/container/fMIndex.h:446-510, 671-754
//   FMIndex::extend_backward           libs/ma/src/container/fMIndex.cpp:21-101
//   FMIndex::bwt_sa / bwt_invPsi       fMIndex.h:329-343, 788-814
//   BinarySeeding::execute / procesInterval / maximallySpanningExtension / smemExtension
//                                       libs/ma/src/module/binarySeeding.cpp:32-178, inc/ma/module/binarySeeding.h:55-452
//   SegmentVector::forEachSeed filter   libs/ma/inc/ma/container/segment.h:316-349
//
// HBM layout: one 64-byte, 64-byte-aligned block per 128 BWT symbols, like the reference
// (4 x u64 cumulative counts + 32 bytes of symbols), so one bwt_occ4 is exactly one 64-byte line fetched with four
// 128-bit loads. Inside the block the symbols are re-laid out as two BIT PLANES (4 x u32 high bits, 4 x u32 low
// bits, symbol j at bit 31 - (j & 31) of word j >> 5) instead of the reference's interleaved 2-bit codes: a rank is
// then 4 words x (prefix mask, 3 x LOP3, 3 x POPC) instead of 8 words with plane extraction — ~3x fewer integer
// instructions per extend_backward, which is what bounds seeding once the loads are in flight. relayout_block()
// converts a reference block; every SAInterval / SA value stays identical (tests + hostsim).
#pragma once
#include "stl_exact.cuh"
#include <stdint.h>

namespace ma
{

struct U4
{
    unsigned int x, y, z, w;
};

struct DevIndex
{
    const U4* bwt; // 4 x U4 per block, PLANE layout (see relayout_block)
    const long long* sa;
    const unsigned char* pac;
    const long long* contig_start;
    const long long* contig_len;
    long long L2[ 5 ];
    long long primary, ref_len, fwd_len;
    int sa_intv, n_contigs;
};

MA_HD inline int popc32( unsigned int v )
{
#if defined( __CUDA_ARCH__ )
    return __popc( v );
#else
    return __builtin_popcount( v );
#endif
}

MA_HD inline U4 ld_u4( const U4* p )
{
#if defined( __CUDA_ARCH__ )
    const uint4 v = __ldg( reinterpret_cast<const uint4*>( p ) );
    return U4{ v.x, v.y, v.z, v.w };
#else
    return *p;
#endif
}

// counts of C, G, T among the first `nvalid` (0..16) symbols of a 16-symbol word of the REFERENCE layout (MSB first)
MA_HD inline void count_word( unsigned int w, int nvalid, int& c, int& g, int& t )
{
    if( nvalid <= 0 )
        return;
    if( nvalid < 16 )
        w &= 0xFFFFFFFFu << ( 32 - 2 * nvalid );
    const unsigned int hi = ( w >> 1 ) & 0x55555555u, lo = w & 0x55555555u;
    t += popc32( hi & lo );
    g += popc32( hi & ~lo );
    c += popc32( lo & ~hi );
}

// reference block (16 x u32: counts, then 8 words of 16 interleaved 2-bit symbols) -> plane block (counts, hi[4], lo[4])
MA_HD inline void relayout_block( const unsigned int* in, unsigned int* out )
{
    for( int i = 0; i < 8; i++ )
        out[ i ] = in[ i ];
    for( int wd = 0; wd < 4; wd++ )
    {
        unsigned int hi = 0, lo = 0;
        for( int j = 0; j < 32; j++ )
        {
            const int sym = wd * 32 + j;
            const unsigned int code = in[ 8 + ( sym >> 4 ) ] >> ( ( ~sym & 15 ) << 1 ) & 3;
            hi |= ( code >> 1 ) << ( 31 - j ), lo |= ( code & 1 ) << ( 31 - j );
        }
        out[ 8 + wd ] = hi, out[ 12 + wd ] = lo;
    }
}

// counts of C, G, T among the first `nvalid` (<= 0 .. >= 32) symbols of one 32-symbol plane pair
MA_HD inline void count_planes( unsigned int hi, unsigned int lo, int nvalid, int& c, int& g, int& t )
{
    const unsigned int m = nvalid >= 32 ? 0xFFFFFFFFu : ( nvalid <= 0 ? 0u : ~( 0xFFFFFFFFu >> nvalid ) );
    t += popc32( hi & lo & m );
    g += popc32( hi & ~lo & m );
    c += popc32( ~hi & lo & m );
}

// bwt_occ4: number of A,C,G,T in BWT[0..k] (k == -1 -> zeros)
MA_HD inline void occ4( const DevIndex& I, long long k, long long cnt[ 4 ] )
{
    if( k == -1 )
    {
        cnt[ 0 ] = cnt[ 1 ] = cnt[ 2 ] = cnt[ 3 ] = 0;
        return;
    }
    k -= ( k >= I.primary );
    const U4* blk = I.bwt + ( ( k >> 7 ) << 2 );
    const U4 c0 = ld_u4( blk ), c1 = ld_u4( blk + 1 ), ph = ld_u4( blk + 2 ), pl = ld_u4( blk + 3 );
    const int n = (int)( k & 127 ) + 1; // symbols of this block to count
    int c = 0, g = 0, t = 0;
    count_planes( ph.x, pl.x, n, c, g, t );
    count_planes( ph.y, pl.y, n - 32, c, g, t );
    count_planes( ph.z, pl.z, n - 64, c, g, t );
    count_planes( ph.w, pl.w, n - 96, c, g, t );
    cnt[ 0 ] = (long long)( ( (unsigned long long)c0.y << 32 ) | c0.x ) + ( n - c - g - t );
    cnt[ 1 ] = (long long)( ( (unsigned long long)c0.w << 32 ) | c0.z ) + c;
    cnt[ 2 ] = (long long)( ( (unsigned long long)c1.y << 32 ) | c1.x ) + g;
    cnt[ 3 ] = (long long)( ( (unsigned long long)c1.w << 32 ) | c1.z ) + t;
}

struct SAI
{
    long long start, rev, size;
};

MA_HD inline SAI sai_rc( const SAI& a )
{
    return SAI{ a.rev, a.start, a.size };
}

MA_HD inline SAI init_interval( const DevIndex& I, int c )
{
    return SAI{ I.L2[ c ] + 1, I.L2[ 3 - c ] + 1, I.L2[ c + 1 ] - I.L2[ c ] };
}

MA_HD inline SAI extend_backward( const DevIndex& I, const SAI& ik, int c )
{
    if( c >= 4 )
        return SAI{ 0, 0, 0 };
    long long ck[ 4 ], cl[ 4 ];
    occ4( I, ik.start - 1, ck );
    occ4( I, ik.start + ik.size - 1, cl );
    const long long s0 = cl[ 0 ] - ck[ 0 ], s1 = cl[ 1 ] - ck[ 1 ], s2 = cl[ 2 ] - ck[ 2 ], s3 = cl[ 3 ] - ck[ 3 ];
    long long k2_0 = ik.rev;
    if( ik.start <= I.primary && ik.start + ik.size > I.primary )
        k2_0++;
    const long long k2_1 = k2_0 + s3, k2_2 = k2_1 + s2, k2_3 = k2_2 + s1; // cntk_2[i] = cntk_2[i-1] + cnts[3-(i-1)]
    switch( c )
    {
        case 0: return SAI{ I.L2[ 0 ] + ck[ 0 ] + 1, k2_3, s0 };
        case 1: return SAI{ I.L2[ 1 ] + ck[ 1 ] + 1, k2_2, s1 };
        case 2: return SAI{ I.L2[ 2 ] + ck[ 2 ] + 1, k2_1, s2 };
        default: return SAI{ I.L2[ 3 ] + ck[ 3 ] + 1, k2_0, s3 };
    }
}

// bwt_invPsi (fMIndex.h:329-343): one block read (B0 and occ hit the same 64-byte block)
MA_HD inline long long inv_psi( const DevIndex& I, long long k )
{
    if( k == I.primary )
        return 0;
    const long long x = k - ( k > I.primary );
    // occ(k, c): k == ref_len -> total count; else inclusive count through k - (k >= primary) == x for k != primary
    const U4* blk = I.bwt + ( ( x >> 7 ) << 2 );
    const U4 c0 = ld_u4( blk ), c1 = ld_u4( blk + 1 ), ph = ld_u4( blk + 2 ), pl = ld_u4( blk + 3 );
    const int pos = (int)( x & 127 );
    const int wsel = pos >> 5, bit = 31 - ( pos & 31 );
    const unsigned int hw = wsel == 0 ? ph.x : wsel == 1 ? ph.y : wsel == 2 ? ph.z : ph.w;
    const unsigned int lw = wsel == 0 ? pl.x : wsel == 1 ? pl.y : wsel == 2 ? pl.z : pl.w;
    const int ch = (int)( ( ( hw >> bit ) & 1 ) << 1 | ( ( lw >> bit ) & 1 ) );
    if( k == I.ref_len )
        return I.L2[ ch ] + ( I.L2[ ch + 1 ] - I.L2[ ch ] );
    const int n = pos + 1;
    int c = 0, g = 0, t = 0;
    count_planes( ph.x, pl.x, n, c, g, t );
    count_planes( ph.y, pl.y, n - 32, c, g, t );
    count_planes( ph.z, pl.z, n - 64, c, g, t );
    count_planes( ph.w, pl.w, n - 96, c, g, t );
    long long occ;
    switch( ch )
    {
        case 0: occ = (long long)( ( (unsigned long long)c0.y << 32 ) | c0.x ) + ( n - c - g - t ); break;
        case 1: occ = (long long)( ( (unsigned long long)c0.w << 32 ) | c0.z ) + c; break;
        case 2: occ = (long long)( ( (unsigned long long)c1.y << 32 ) | c1.x ) + g; break;
        default: occ = (long long)( ( (unsigned long long)c1.w << 32 ) | c1.z ) + t; break;
    }
    return I.L2[ ch ] + occ;
}

// bwt_sa (fMIndex.h:788-814); *pSteps receives the number of invPsi steps (roofline accounting)
MA_HD inline long long bwt_sa( const DevIndex& I, long long k, int* pSteps )
{
    long long s = 0;
    const long long mask = I.sa_intv - 1;
    while( k & mask )
    {
        ++s;
        k = inv_psi( I, k );
    }
    if( pSteps )
        *pSteps = (int)s;
    return s + I.sa[ k / I.sa_intv ];
}

// ---- pack --------------------------------------------------------------------------------------------------
MA_HD inline int pack_nuc( const DevIndex& I, long long pos )
{
    return I.pac[ pos >> 2 ] >> ( ( ~pos & 3 ) << 1 ) & 3;
}
// base of the virtual text forward ++ reverse-complement at position p (Pack::vExtract, pack.h:1429-1435)
MA_HD inline int pack_virtual( const DevIndex& I, long long p )
{
    return p < I.fwd_len ? pack_nuc( I, p ) : 3 - pack_nuc( I, 2 * I.fwd_len - 1 - p );
}
// Pack::uiSequenceIdForPosition (pack.h:933-990) including the fall-through of its binary search
MA_HD inline long long seq_id_for_position( const DevIndex& I, long long pos )
{
    const long long a = pos >= I.fwd_len ? 2 * I.fwd_len - ( pos + 1 ) : pos;
    unsigned long long l = 0, m = 0, r = (unsigned long long)I.n_contigs;
    while( l < r )
    {
        m = ( l + r ) / 2;
        if( a >= I.contig_start[ m ] )
        {
            if( m == (unsigned long long)I.n_contigs - 1 )
                break;
            if( a < I.contig_start[ m + 1 ] )
                break;
            l = m + 1;
        }
        else
            r = m;
    }
    return (long long)m;
}
MA_HD inline long long seq_id_or_rev( const DevIndex& I, long long pos ) // pack.h:1029-1034
{
    if( pos >= I.fwd_len )
        return seq_id_for_position( I, 2 * I.fwd_len - ( pos + 1 ) ) * 2 + 1;
    return seq_id_for_position( I, pos ) * 2;
}
MA_HD inline long long end_of_seq_or_rev( const DevIndex& I, long long id ) // pack.h:1040-1045
{
    if( id % 2 == 1 )
        return ( 2 * I.fwd_len - ( I.contig_start[ id / 2 ] + 1 ) ) - 1;
    return I.contig_start[ id / 2 ] + I.contig_len[ id / 2 ];
}
MA_HD inline long long start_of_seq_or_rev( const DevIndex& I, long long id ) // pack.h:1047-1052
{
    if( id % 2 == 1 )
        return ( 2 * I.fwd_len - ( I.contig_start[ id / 2 ] + I.contig_len[ id / 2 ] + 1 ) ) + 1;
    return I.contig_start[ id / 2 ];
}
MA_HD inline bool bridging_subsection( const DevIndex& I, long long begin, long long size ) // pack.h:1072-1087
{
    if( size <= 0 )
        return false;
    return ( begin >= I.fwd_len ) != ( begin + size - 1 >= I.fwd_len ) ||
           seq_id_or_rev( I, begin ) != seq_id_or_rev( I, begin + size - 1 );
}

// ---- seeding -------------------------------------------------------------------------------------------------
struct SeedParams
{
    int technique; // 0 maxSpan, 1 SMEM
    int min_amb, max_amb;
    int min_seed_len;
    int drop_min_size;
    double drop_factor;
    int disable_heuristics;
    long long genome_size_disable;
};

struct SegRec // one segment (query interval + SA interval); size = length - 1
{
    int start, size;
    SAI sa;
};

// Per-thread working memory for the SMEM interval lists (two ping-pong lists of up to cap entries)
struct SeedScratch
{
    SegRec* listA;
    SegRec* listB;
    int cap;
};

// The ONE interval list of the state-machine seeder (SeederSM). The reference keeps two lists per SMEM centre and
// ping-pongs between them (binarySeeding.h smemExtension); the backward pass reads entry j and appends at most one
// entry at a position <= j, so it can run in place, and every entry of a pass shares the same query start, which is
// therefore kept as a scalar. What is left per entry is (length - 1, k, l, s): packed here into 16 + 4 bytes (three
// 40-bit values, reference length < 2^40) so that the first K entries of every thread fit in shared memory
// (entry j of thread t at [ j * stride + t ]: bank = t mod 8 for the 16-byte part, conflict free for any mix of j).
// Entries >= K (long lists: repetitive reads) spill to the thread's global scratch.
// Every entry also carries a multiplicity: the number of reference list entries it stands for (SeederSM::consume_bwd).
struct SegList
{
    U4* pk; // 3 x 40 bit
    int* sz;
    unsigned short* mu; // multiplicity (see SeederSM::consume_bwd)
    int stride, K;
    SegRec* ovf;
    int cap; // total capacity (K + overflow entries)

    MA_HD static U4 pack( const SAI& a )
    {
        U4 v;
        v.x = (unsigned int)a.start, v.y = (unsigned int)a.rev, v.z = (unsigned int)a.size;
        v.w = (unsigned int)( ( a.start >> 32 ) & 0xFF ) | (unsigned int)( ( a.rev >> 32 ) & 0xFF ) << 8 |
              (unsigned int)( ( a.size >> 32 ) & 0xFF ) << 16;
        return v;
    }
    MA_HD static SAI unpack( const U4& v )
    {
        SAI a;
        a.start = (long long)v.x | (long long)( v.w & 0xFF ) << 32;
        a.rev = (long long)v.y | (long long)( ( v.w >> 8 ) & 0xFF ) << 32;
        a.size = (long long)v.z | (long long)( ( v.w >> 16 ) & 0xFF ) << 32;
        return a;
    }
    MA_HD SAI sa( int j ) const
    {
        return j < K ? unpack( pk[ j * stride ] ) : ovf[ j - K ].sa;
    }
    MA_HD int size( int j ) const
    {
        return j < K ? sz[ j * stride ] : ovf[ j - K ].size;
    }
    MA_HD int mult( int j ) const
    {
        return j < K ? (int)mu[ j * stride ] : ovf[ j - K ].start;
    }
    MA_HD void add_mult( int j, int m )
    {
        if( j < K )
            mu[ j * stride ] = (unsigned short)( mu[ j * stride ] + m );
        else
            ovf[ j - K ].start += m;
    }
    MA_HD void set( int j, int size, const SAI& a, int m = 1 )
    {
        if( j < K )
            pk[ j * stride ] = pack( a ), sz[ j * stride ] = size, mu[ j * stride ] = (unsigned short)m;
        else
            ovf[ j - K ] = SegRec{ m, size, a };
    }
};

// Sink interface: void seg( const SegRec& ) is called for every segment in the reference's emission order
// (DFS pre-order).  Returns false in *pOverflow if a list overflowed its capacity.
template <class Sink> struct Seeder
{
    const DevIndex& I;
    const SeedParams& P;
    const unsigned char* q;
    const int L;
    SeedScratch S;
    Sink& sink;
    long long nExt = 0;
    bool overflow = false;
    int lastStart = -1, lastEnd = -1; // start/end() of the most recently emitted segment (maxSpan duplicate check)

    MA_HD Seeder( const DevIndex& I, const SeedParams& P, const unsigned char* q, int L, SeedScratch S, Sink& sink )
        : I( I ), P( P ), q( q ), L( L ), S( S ), sink( sink )
    {}
    MA_HD static int comp( int c )
    {
        return c < 4 ? 3 - c : 5;
    }
    MA_HD SAI ext( const SAI& ik, int c )
    {
        nExt++;
        return extend_backward( I, ik, c );
    }
    MA_HD bool stop( const SAI& ok, const SAI& ik ) const
    {
        return ok.size <= 0 || ( ok.size <= P.min_amb && ik.size <= P.max_amb );
    }
    MA_HD void emit( int start, int size, const SAI& sa )
    {
        SegRec r{ start, size, sa };
        lastStart = start, lastEnd = start + size;
        sink.seg( r );
    }
    // binarySeeding.h:55-252. cs/ce = covered start / end() (index of the last covered base, or center+1 for N)
    MA_HD void maxSpan( int center, int& cs, int& ce )
    {
        if( q[ center ] >= 4 )
        {
            cs = center, ce = center + 1;
            return;
        }
        SAI ik = init_interval( I, comp( q[ center ] ) );
        if( ik.size == 0 )
        {
            cs = center, ce = center + 1;
            return;
        }
        int end = center;
        for( int i = center + 1; i < L; i++ )
        {
            SAI ok = ext( ik, comp( q[ i ] ) );
            if( stop( ok, ik ) )
                break;
            end = i, ik = ok;
        }
        ik = sai_rc( ik );
        int start = center;
        for( int i = center - 1; i >= 0; i-- )
        {
            SAI ok = ext( ik, q[ i ] );
            if( stop( ok, ik ) )
                break;
            start = i, ik = ok;
        }
        emit( start, end - start, ik );
        const int s1 = start, e1 = end;
        ik = init_interval( I, q[ center ] );
        start = center;
        for( int i = center - 1; i >= 0; i-- )
        {
            SAI ok = ext( ik, q[ i ] );
            if( stop( ok, ik ) )
                break;
            start = i, ik = ok;
        }
        ik = sai_rc( ik );
        end = center;
        for( int i = center + 1; i < L; i++ )
        {
            SAI ok = ext( ik, comp( q[ i ] ) );
            if( stop( ok, ik ) )
                break;
            end = i, ik = ok;
        }
        if( s1 == start && e1 == end )
        {
            cs = s1, ce = e1;
            return;
        }
        emit( start, end - start, sai_rc( ik ) );
        cs = s1 < start ? s1 : start;
        ce = e1 > end ? e1 : end;
    }
    // binarySeeding.h:261-452
    MA_HD void smem( int center, int& cs, int& ce )
    {
        cs = center, ce = center;
        if( q[ center ] >= 4 )
        {
            ce = center + 1;
            return;
        }
        SAI ik = init_interval( I, comp( q[ center ] ) );
        SegRec* curr = S.listA;
        SegRec* next = S.listB;
        int nCurr = 0, nNext = 0;
        for( int i = center + 1; i < L; i++ )
        {
            SAI ok = ext( ik, comp( q[ i ] ) );
            if( ok.size != ik.size )
            {
                if( nCurr < S.cap )
                    curr[ nCurr ] = SegRec{ center, i - center - 1, sai_rc( ik ) };
                else
                    overflow = true;
                nCurr++;
            }
            if( i == L - 1 && ok.size != 0 )
            {
                if( nCurr < S.cap )
                    curr[ nCurr ] = SegRec{ center, i - center, sai_rc( ok ) };
                else
                    overflow = true;
                nCurr++;
            }
            if( ok.size == 0 )
                break;
            if( ok.size <= P.min_amb && ik.size <= P.max_amb )
                break;
            ik = ok;
            ce = i;
        }
        if( nCurr > S.cap )
            nCurr = S.cap;
        for( int a = 0, b = nCurr - 1; a < b; a++, b-- ) // std::reverse
            stl::swp( curr[ a ], curr[ b ] );
        if( center != 0 )
            for( int i = center - 1; i >= 0; i-- )
            {
                bool bHaveOne = false;
                nNext = 0;
                for( int j = 0; j < nCurr; j++ )
                {
                    const SegRec s = curr[ j ];
                    SAI ok = ext( s.sa, q[ i ] );
                    if( ok.size <= P.min_amb && !bHaveOne )
                    {
                        emit( s.start, s.size, s.sa );
                        bHaveOne = true;
                    }
                    // sic: s.size is the query-interval size field (binarySeeding.h:404)
                    else if( ok.size > P.min_amb || ( ok.size > 0 && s.size >= P.max_amb ) )
                        next[ nNext++ ] = SegRec{ i, s.size + 1, ok };
                }
                SegRec* t = curr;
                curr = next, next = t;
                nCurr = nNext;
                if( nCurr == 0 )
                    break;
                cs = i;
                if( i == 0 )
                    break;
            }
        if( nCurr != 0 )
            emit( curr[ 0 ].start, curr[ 0 ].size, curr[ 0 ].sa );
    }
    // binarySeeding.cpp:32-84 with the recursion replaced by an explicit LIFO of pending right-hand areas
    MA_HD void run( )
    {
        if( L <= 0 )
            return;
        const int STK = 24; // depth <= log2(L) + 2
        int stS[ STK ], stN[ STK ];
        int sp = 0;
        stS[ sp ] = 0, stN[ sp ] = L, sp++;
        while( sp > 0 )
        {
            --sp;
            const int aStart = stS[ sp ], aSize = stN[ sp ];
            const int center = aStart + aSize / 2;
            int cs, ce;
            if( P.technique == 0 )
                maxSpan( center, cs, ce );
            else
                smem( center, cs, ce );
            const int aEnd = aStart + aSize;
            // the right remainder is processed after the complete left subtree: push it first
            if( aEnd > ce + 1 )
            {
                if( sp < STK )
                    stS[ sp ] = ce, stN[ sp ] = aEnd - ce, sp++;
                else
                    overflow = true;
            }
            if( cs != 0 && aStart + 1 < cs )
            {
                if( sp < STK )
                    stS[ sp ] = aStart, stN[ sp ] = cs - aStart, sp++;
                else
                    overflow = true;
            }
        }
    }
};


// Same algorithm as Seeder, re-expressed as a resumable state machine with ONE extend_backward call site.
// A warp holds 32 reads in different phases (forward sweep, backward sweep over an interval list, ...): with the
// recursive formulation the lanes diverge and only a few of them have their two 64-byte occ-block loads in flight
// at any time. Here every lane of the warp reaches the same call site each iteration, so all 64 loads of the warp
// are issued together; only the cheap bookkeeping around it diverges.
template <class Sink> struct SeederSM
{
    enum Phase
    {
        P_NEXT_CENTER,
        P_SMEM_FWD,
        P_SMEM_BWD,
        P_MS1_FWD,
        P_MS1_BWD,
        P_MS2_BWD,
        P_MS2_FWD,
        P_DONE
    };
    const DevIndex& I;
    const SeedParams& P;
    const unsigned char* q;
    int L;
    SegList S;
    Sink& sink;
    long long nExt = 0;
    bool overflow = false;
    // interval stack (depth <= log2(L) + 2) in caller-provided memory: keeps this struct array-free (registers)
    int* stS;
    int* stN;
    int sp = 0;
    // state of the centre being processed
    int phase = P_NEXT_CENTER;
    int aStart = 0, aSize = 0, center = 0, cs = 0, ce = 0, i = 0;
    SAI ik;
    int start = 0, end = 0, s1 = 0, e1 = 0;
    int lstart = 0; // query start shared by all entries of the list
    int nCurr = 0, nNext = 0, j = 0;
    bool bHaveOne = false;
    // last entry pushed in the running backward pass (its SA interval), for the merge of identical intervals
    long long lastStart = 0;
    int lastSize = -1; // < 0: none (intervals of 2^31 rows or more are never merged)
    unsigned int nLookup = 0; // extend_backward calls really made (nExt counts the reference's)

    MA_HD SeederSM( const DevIndex& I, const SeedParams& P, const unsigned char* q, int L, SegList S, Sink& sink,
                    int* pStack /* 2 x 40 ints */ )
        : I( I ), P( P ), q( q ), L( L ), S( S ), sink( sink ), stS( pStack ), stN( pStack + 40 )
    {
        ik = SAI{ 0, 0, 0 };
        begin( q, L );
    }
    // (re)starts the machine on a new read
    MA_HD void begin( const unsigned char* q_, int L_ )
    {
        q = q_, L = L_;
        nExt = 0, nLookup = 0, overflow = false, sp = 0;
        phase = P_NEXT_CENTER;
        if( L > 0 )
            stS[ 0 ] = 0, stN[ 0 ] = L, sp = 1;
        else
            phase = P_DONE;
    }
    MA_HD static int comp( int c )
    {
        return c < 4 ? 3 - c : 5;
    }
    MA_HD bool stop( const SAI& ok, const SAI& ikk ) const
    {
        return ok.size <= 0 || ( ok.size <= P.min_amb && ikk.size <= P.max_amb );
    }
    MA_HD void emit( int st, int size, const SAI& sa )
    {
        sink.seg( SegRec{ st, size, sa } );
    }
    MA_HD void push_list( int& n, int size, const SAI& sa, int m = 1 )
    {
        if( n < S.cap )
            S.set( n, size, sa, m );
        else
            overflow = true;
        n++;
    }
    // binarySeeding.cpp:58-83 after the extension of one centre
    MA_HD void finish_center( )
    {
        const int aEnd = aStart + aSize;
        if( aEnd > ce + 1 )
        {
            if( sp < 40 )
                stS[ sp ] = ce, stN[ sp ] = aEnd - ce, sp++;
            else
                overflow = true;
        }
        if( cs != 0 && aStart + 1 < cs )
        {
            if( sp < 40 )
                stS[ sp ] = aStart, stN[ sp ] = cs - aStart, sp++;
            else
                overflow = true;
        }
        phase = P_NEXT_CENTER;
    }
    MA_HD void smem_fwd_done( )
    {
        if( nCurr > S.cap )
            nCurr = S.cap;
        for( int a = 0, b = nCurr - 1; a < b; a++, b-- )
        {
            const int sa_ = S.size( a ), sb_ = S.size( b );
            const SAI ia = S.sa( a ), ib = S.sa( b );
            S.set( a, sb_, ib ), S.set( b, sa_, ia ); // forward-pass entries all have multiplicity 1
        }
        lstart = center;
        if( center != 0 )
        {
            i = center - 1, j = 0, bHaveOne = false, nNext = 0, lastSize = -1;
            phase = P_SMEM_BWD;
        }
        else
            smem_final( );
    }
    MA_HD void smem_final( )
    {
        if( nCurr != 0 )
            emit( lstart, S.size( 0 ), S.sa( 0 ) );
        finish_center( );
    }
    MA_HD void ms_emit_first( )
    {
        emit( start, end - start, ik );
        s1 = start, e1 = end;
        ik = init_interval( I, q[ center ] );
        start = center;
        i = center - 1;
        phase = P_MS2_BWD;
    }
    MA_HD void ms_finish( )
    {
        if( s1 == start && e1 == end )
            cs = s1, ce = e1;
        else
        {
            emit( start, end - start, sai_rc( ik ) );
            cs = s1 < start ? s1 : start;
            ce = e1 > end ? e1 : end;
        }
        finish_center( );
    }
    // Advances the bookkeeping until the next extension is needed. Returns false when the read is finished.
    MA_HD bool request( SAI& rIk, int& rC )
    {
        while( true )
        {
            switch( phase )
            {
                case P_NEXT_CENTER:
                {
                    if( sp == 0 )
                    {
                        phase = P_DONE;
                        return false;
                    }
                    --sp;
                    aStart = stS[ sp ], aSize = stN[ sp ];
                    center = aStart + aSize / 2;
                    if( q[ center ] >= 4 )
                    {
                        cs = center, ce = center + 1;
                        finish_center( );
                        break;
                    }
                    ik = init_interval( I, comp( q[ center ] ) );
                    if( P.technique == 0 )
                    {
                        if( ik.size == 0 )
                        {
                            cs = center, ce = center + 1;
                            finish_center( );
                            break;
                        }
                        end = center, i = center + 1;
                        phase = P_MS1_FWD;
                    }
                    else
                    {
                        cs = center, ce = center;
                        nCurr = 0, nNext = 0;
                        i = center + 1;
                        phase = P_SMEM_FWD;
                    }
                    break;
                }
                case P_SMEM_FWD:
                    if( i < L )
                    {
                        rIk = ik, rC = comp( q[ i ] );
                        return true;
                    }
                    smem_fwd_done( );
                    break;
                case P_SMEM_BWD:
                    if( j < nCurr )
                    {
                        rIk = S.sa( j ), rC = q[ i ];
                        return true;
                    }
                    { // the pass over position i is complete: the entries written in place are the new list
                        nCurr = nNext < S.cap ? nNext : S.cap;
                        lstart = i;
                        if( nCurr == 0 )
                        {
                            finish_center( ); // nothing left to emit
                            break;
                        }
                        cs = i;
                        if( i == 0 )
                        {
                            smem_final( );
                            break;
                        }
                        i--, j = 0, bHaveOne = false, nNext = 0, lastSize = -1;
                    }
                    break;
                case P_MS1_FWD:
                    if( i < L )
                    {
                        rIk = ik, rC = comp( q[ i ] );
                        return true;
                    }
                    ik = sai_rc( ik ), start = center, i = center - 1, phase = P_MS1_BWD;
                    break;
                case P_MS1_BWD:
                    if( i >= 0 )
                    {
                        rIk = ik, rC = q[ i ];
                        return true;
                    }
                    ms_emit_first( );
                    break;
                case P_MS2_BWD:
                    if( i >= 0 )
                    {
                        rIk = ik, rC = q[ i ];
                        return true;
                    }
                    ik = sai_rc( ik ), end = center, i = center + 1, phase = P_MS2_FWD;
                    break;
                case P_MS2_FWD:
                    if( i < L )
                    {
                        rIk = ik, rC = comp( q[ i ] );
                        return true;
                    }
                    ms_finish( );
                    break;
                default:
                    return false;
            }
        }
    }
    // Feeds the result of the requested extension back into the state machine.
    MA_HD void consume( const SAI& ok )
    {
        nExt++, nLookup++;
        switch( phase )
        {
            case P_SMEM_FWD:
            {
                if( ok.size != ik.size )
                    push_list( nCurr, i - center - 1, sai_rc( ik ) );
                if( i == L - 1 && ok.size != 0 )
                    push_list( nCurr, i - center, sai_rc( ok ) );
                if( ok.size == 0 || ( ok.size <= P.min_amb && ik.size <= P.max_amb ) )
                {
                    smem_fwd_done( );
                    break;
                }
                ik = ok;
                ce = i;
                i++;
                break;
            }
            case P_SMEM_BWD:
            {
                // Entries with the same SA interval (start, size) extend identically for ever and, with
                // min_amb == 0, only the first of them (the longest match) can ever be emitted: the failing entries of a
                // pass are a prefix of the list and just the first failing one is reported (bHaveOne). Such entries
                // are therefore merged into the first one, which keeps count of how many reference entries it stands
                // for so that nExt stays the reference's number of extend_backward calls.
                const int sSize = S.size( j ), m = S.mult( j );
                nExt += m - 1;
                if( ok.size <= P.min_amb && !bHaveOne )
                {
                    emit( lstart, sSize, S.sa( j ) );
                    bHaveOne = true;
                }
                else if( ok.size > P.min_amb || ( ok.size > 0 && sSize >= P.max_amb ) )
                {
                    if( P.min_amb == 0 && ok.start == lastStart && ok.size == (long long)lastSize )
                        S.add_mult( nNext - 1, m );
                    else
                    {
                        push_list( nNext, sSize + 1, ok, m ); // in place: nNext <= j, entry j has been consumed
                        lastStart = ok.start, lastSize = ok.size < 0x7fffffffll ? (int)ok.size : -1;
                    }
                }
                j++;
                break;
            }
            case P_MS1_FWD:
                if( stop( ok, ik ) )
                    ik = sai_rc( ik ), start = center, i = center - 1, phase = P_MS1_BWD;
                else
                    end = i, ik = ok, i++;
                break;
            case P_MS1_BWD:
                if( stop( ok, ik ) )
                    ms_emit_first( );
                else
                    start = i, ik = ok, i--;
                break;
            case P_MS2_BWD:
                if( stop( ok, ik ) )
                    ik = sai_rc( ik ), end = center, i = center + 1, phase = P_MS2_FWD;
                else
                    start = i, ik = ok, i--;
                break;
            case P_MS2_FWD:
                if( stop( ok, ik ) )
                    ms_finish( );
                else
                    end = i, ik = ok, i++;
                break;
            default:
                break;
        }
    }
    MA_HD void run( )
    {
        SAI rIk;
        int rC;
        while( request( rIk, rC ) )
            consume( extend_backward( I, rIk, rC ) );
    }
};

} // namespace ma
