// CPU ORACLE — test infrastructure only (see oracle.h).
// FM-index accessors, BinarySeeding (maxSpan + SMEM) and ExtractSeeds, restated from
//   libs/ma/inc/ma/container/fMIndex.h:195-230, 329-343, 434-510, 555-663, 768-814   (FMIndex)
//   libs/ma/src/container/fMIndex.cpp:21-101                                           (extend_backward)
//   libs/ma/src/module/binarySeeding.cpp:32-178, libs/ma/inc/ma/module/binarySeeding.h:55-452
//   libs/ma/inc/ma/container/segment.h:89-113, 316-369, libs/ma/inc/ma/module/stripOfConsideration.h:42-157
//   libs/ma/inc/ma/container/pack.h:172-176, 275-451, 933-1087, 1147-1236               (Pack)
#include "ma_oracle.h"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace oracle
{

bool Params::preset( std::string s )
{
    std::string t;
    for( char c : s )
        if( c != '_' && c != ' ' && c != '-' )
            t += (char)tolower( c );
    *this = Params( );
    if( t == "default" )
        return true;
    if( t == "illumina" || t == "illuminapaired" ) // parameter.h:1083-1094
    {
        seeding_technique = 1, max_ambiguity = 500, min_num_soc = 10, max_num_soc = 20;
        use_paired_reads = t == "illuminapaired";
        return true;
    }
    if( t == "pacbio" ) // parameter.h:1096-1098
    {
        max_supplementary_per_prim = 100;
        min_num_soc = 5;
        return true;
    }
    if( t == "nanopore" ) // parameter.h:1101-1104
    {
        seeding_technique = 1, min_num_soc = 5;
        max_supplementary_per_prim = 100;
        return true;
    }
    return false;
}

static std::vector<char> readFile( const std::string& f )
{
    std::ifstream in( f, std::ios::binary );
    if( !in )
        throw std::runtime_error( "cannot open " + f );
    return std::vector<char>( ( std::istreambuf_iterator<char>( in ) ), std::istreambuf_iterator<char>( ) );
}

void Index::load( const std::string& p )
{
    { // .bwt: primary, L2[1..4], words (fMIndex.h:555-599)
        auto b = readFile( p + ".bwt" );
        memcpy( &primary, b.data( ), 8 );
        memcpy( &L2[ 1 ], b.data( ) + 8, 32 );
        L2[ 0 ] = 0;
        bwt.resize( ( b.size( ) - 40 ) / 4 );
        memcpy( bwt.data( ), b.data( ) + 40, bwt.size( ) * 4 );
        ref_len = (int64_t)L2[ 4 ];
    }
    { // .sa: primary, L2[1..4], sa_intv (int), seq_len, sa[1..] (fMIndex.h:602-663)
        auto b = readFile( p + ".sa" );
        memcpy( &sa_intv, b.data( ) + 40, 4 );
        size_t n = ( ref_len + sa_intv ) / sa_intv;
        sa.assign( n, 0 );
        sa[ 0 ] = -1;
        memcpy( &sa[ 1 ], b.data( ) + 52, ( n - 1 ) * 8 );
    }
    { // .ann (pack.h:230-269, 353-451)
        std::ifstream in( p + ".ann" );
        if( !in )
            throw std::runtime_error( "cannot open " + p + ".ann" );
        int64_t nSeq, seed;
        in >> fwd_len >> nSeq >> seed;
        std::string line;
        std::getline( in, line );
        for( int64_t i = 0; i < nSeq; i++ )
        {
            Contig c;
            std::getline( in, line );
            std::istringstream ls( line );
            int64_t gi;
            ls >> gi >> c.name;
            int64_t holes;
            in >> c.start >> c.length >> holes;
            std::getline( in, line );
            contigs.push_back( c );
        }
    }
    {
        auto b = readFile( p + ".pac" );
        pac.assign( b.begin( ), b.begin( ) + ( fwd_len + 3 ) / 4 );
    }
}

static inline uint32_t occ_aux4( uint32_t w ) // fMIndex.h:421-427 (byte LUT replaced by arithmetic, same values)
{
    uint32_t x = 0;
    for( int j = 0; j < 16; j++ )
        x += 1u << ( ( ( w >> ( 2 * j ) ) & 3 ) * 8 );
    return x;
}

void Index::occ4( int64_t k, int64_t cnt[ 4 ] ) const // fMIndex.h:446-510
{
    if( k == -1 )
    {
        cnt[ 0 ] = cnt[ 1 ] = cnt[ 2 ] = cnt[ 3 ] = 0;
        return;
    }
    k -= ( k >= primary );
    const uint32_t* p = &bwt[ ( k >> 7 ) << 4 ];
    memcpy( cnt, p, 32 );
    p += 8;
    const uint32_t* end = p + ( ( k >> 4 ) - ( ( k & ~127ll ) >> 4 ) );
    uint64_t x = 0;
    for( ; p < end; ++p )
        x += occ_aux4( *p );
    uint32_t tmp = *p & ~( ( 1U << ( ( ~k & 15 ) << 1 ) ) - 1 );
    x += occ_aux4( tmp ) - ( ~k & 15 );
    cnt[ 0 ] += x & 0xff, cnt[ 1 ] += x >> 8 & 0xff, cnt[ 2 ] += x >> 16 & 0xff, cnt[ 3 ] += x >> 24 & 0xff;
}

int64_t Index::occ( int64_t k, int c ) const // fMIndex.h:287-325
{
    if( k == ref_len )
        return L2[ c + 1 ] - L2[ c ];
    if( k == -1 )
        return 0;
    int64_t cnt[ 4 ];
    occ4( k, cnt ); // same inclusive count, one symbol of the four
    return cnt[ c ];
}

int Index::B0( int64_t k ) const // fMIndex.h:255-256
{
    return bwt[ ( ( k >> 7 ) << 4 ) + 8 + ( ( k & 0x7f ) >> 4 ) ] >> ( ( ~k & 0xf ) << 1 ) & 3;
}

int64_t Index::invPsi( int64_t k ) const // fMIndex.h:329-343
{
    int64_t x = k - ( k > primary );
    int c = B0( x );
    x = L2[ c ] + occ( k, c );
    return k == primary ? 0 : x;
}

int64_t Index::bwt_sa( int64_t k ) const // fMIndex.h:788-814
{
    int64_t s = 0;
    while( k & ( sa_intv - 1 ) )
        ++s, k = invPsi( k );
    return s + sa[ k / sa_intv ];
}

SAInterval init_interval( const Index& I, int c ) // fMIndex.h:768-775
{
    return SAInterval{ (int64_t)I.L2[ c ] + 1, (int64_t)I.L2[ 3 - c ] + 1, (int64_t)( I.L2[ c + 1 ] - I.L2[ c ] ) };
}

SAInterval extend_backward( const Index& I, const SAInterval& ik, int c ) // fMIndex.cpp:21-101
{
    if( c >= 4 )
        return SAInterval{ 0, 0, 0 };
    int64_t cntk[ 4 ], cntl[ 4 ], cnts[ 4 ], cntk_2[ 4 ];
    I.occ4( ik.start - 1, cntk );
    I.occ4( ik.end( ) - 1, cntl );
    for( int i = 0; i < 4; i++ )
        cnts[ i ] = cntl[ i ] - cntk[ i ];
    cntk_2[ 0 ] = ik.rev;
    if( ik.start <= I.primary && ik.end( ) > I.primary )
        cntk_2[ 0 ]++;
    for( int i = 1; i < 4; i++ )
        cntk_2[ i ] = cntk_2[ i - 1 ] + cnts[ 3 - ( i - 1 ) ];
    return SAInterval{ (int64_t)I.L2[ c ] + cntk[ c ] + 1, cntk_2[ 3 - c ], cnts[ c ] };
}

static inline int comp( int c ) // nucSeq.h:524-532
{
    return c < 4 ? 3 - c : 5;
}

namespace
{
struct Seeder
{
    const Index& I;
    const Params& P;
    const std::vector<uint8_t>& q;
    std::vector<Segment> out;
    int64_t nExt = 0;
    const int64_t L;
    Seeder( const Index& I, const Params& P, const std::vector<uint8_t>& q ) : I( I ), P( P ), q( q ), L( q.size( ) )
    {}
    SAInterval ext( const SAInterval& ik, int c )
    {
        nExt++;
        return extend_backward( I, ik, c );
    }
    bool stop( const SAInterval& ok, const SAInterval& ik ) const
    {
        return ok.size <= 0 || ( ok.size <= P.min_ambiguity && ik.size <= P.max_ambiguity );
    }
    // binarySeeding.h:55-252; returns covered (start, end) with end = index of the last covered base
    std::pair<int64_t, int64_t> maxSpan( int64_t center )
    {
        if( q[ center ] >= 4 )
            return { center, center + 1 };
        SAInterval ik = init_interval( I, comp( q[ center ] ) );
        if( ik.size == 0 )
            return { center, center + 1 };
        int64_t end = center;
        for( int64_t i = center + 1; i < L; i++ )
        {
            SAInterval ok = ext( ik, comp( q[ i ] ) );
            if( stop( ok, ik ) )
                break;
            end = i, ik = ok;
        }
        ik = ik.revComp( );
        int64_t start = center;
        for( int64_t i = center - 1; i >= 0; i-- )
        {
            SAInterval ok = ext( ik, q[ i ] );
            if( stop( ok, ik ) )
                break;
            start = i, ik = ok;
        }
        out.push_back( Segment{ start, end - start, ik } );
        const int64_t s1 = start, e1 = end;
        ik = init_interval( I, q[ center ] );
        start = center;
        for( int64_t i = center - 1; i >= 0; i-- )
        {
            SAInterval ok = ext( ik, q[ i ] );
            if( stop( ok, ik ) )
                break;
            start = i, ik = ok;
        }
        ik = ik.revComp( );
        end = center;
        for( int64_t i = center + 1; i < L; i++ )
        {
            SAInterval ok = ext( ik, comp( q[ i ] ) );
            if( stop( ok, ik ) )
                break;
            end = i, ik = ok;
        }
        if( s1 == start && e1 == end )
            return { s1, e1 };
        out.push_back( Segment{ start, end - start, ik.revComp( ) } );
        return { std::min( s1, start ), std::max( e1, end ) };
    }
    // binarySeeding.h:261-452
    std::pair<int64_t, int64_t> smem( int64_t center )
    {
        int64_t rs = center, re = center; // ret(center, 0): start, end()
        if( q[ center ] >= 4 )
            return { center, center + 1 };
        SAInterval ik = init_interval( I, comp( q[ center ] ) );
        std::vector<Segment> curr, next;
        for( int64_t i = center + 1; i < L; i++ )
        {
            SAInterval ok = ext( ik, comp( q[ i ] ) );
            if( ok.size != ik.size )
                curr.push_back( Segment{ center, i - center - 1, ik.revComp( ) } );
            if( i == L - 1 && ok.size != 0 )
                curr.push_back( Segment{ center, i - center, ok.revComp( ) } );
            if( ok.size == 0 )
                break;
            if( ok.size <= P.min_ambiguity && ik.size <= P.max_ambiguity )
                break;
            ik = ok;
            re = i;
        }
        std::reverse( curr.begin( ), curr.end( ) );
        std::vector<Segment>*pPrev = &curr, *pCurr = &next;
        if( center != 0 )
            for( int64_t i = center - 1; i >= 0; i-- )
            {
                bool bHaveOne = false;
                for( const Segment& s : *pPrev )
                {
                    SAInterval ok = ext( s.sa, q[ i ] );
                    if( ok.size <= P.min_ambiguity && !bHaveOne )
                    {
                        out.push_back( s );
                        bHaveOne = true;
                    }
                    // sic: s.size is the QUERY-interval size field, not the SA size (binarySeeding.h:404)
                    else if( ok.size > P.min_ambiguity || ( ok.size > 0 && s.size >= P.max_ambiguity ) )
                        pCurr->push_back( Segment{ i, s.size + 1, ok } );
                }
                std::swap( pPrev, pCurr );
                pCurr->clear( );
                if( pPrev->empty( ) )
                    break;
                rs = i;
                if( i == 0 )
                    break;
            }
        if( !pPrev->empty( ) )
            out.push_back( pPrev->front( ) );
        return { rs, re };
    }
    // binarySeeding.cpp:32-84
    void process( int64_t aStart, int64_t aSize )
    {
        while( true )
        {
            const int64_t center = aStart + aSize / 2;
            auto cov = P.seeding_technique == 0 ? maxSpan( center ) : smem( center );
            if( cov.first != 0 && aStart + 1 < cov.first )
                process( aStart, cov.first - aStart );
            const int64_t aEnd = aStart + aSize;
            if( aEnd > cov.second + 1 )
            {
                aStart = cov.second;
                aSize = aEnd - cov.second;
            }
            else
                break;
        }
    }
};
} // namespace

std::vector<Segment> binary_seeding( const Index& I, const Params& P, const std::vector<uint8_t>& q, int64_t* pnExt )
{
    Seeder S( I, P, q );
    if( q.empty( ) )
        return { };
    S.process( 0, (int64_t)q.size( ) );
    if( pnExt )
        *pnExt = S.nExt;
    // drop-off heuristic, binarySeeding.cpp:172-175, segment.h:278-289
    if( !P.disable_heuristics && P.seed_drop_min_size != 0 )
    {
        size_t sum = 0;
        for( auto& s : S.out )
            sum += (size_t)s.size / (size_t)P.seed_drop_min_size;
        if( (double)sum < P.seed_drop_factor * (double)q.size( ) && (uint64_t)P.genome_size_disable < (uint64_t)I.ref_len )
            S.out.clear( );
    }
    return S.out;
}

int64_t Index::seqIdForPosition( int64_t pos ) const // pack.h:933-990 (binary search incl. its fall-through)
{
    const int64_t a = onReverse( pos ) ? 2 * fwd_len - ( pos + 1 ) : pos;
    uint64_t l = 0, m = 0, r = contigs.size( );
    while( l < r )
    {
        m = ( l + r ) / 2;
        if( a >= contigs[ m ].start )
        {
            if( m == contigs.size( ) - 1 )
                break;
            if( a < contigs[ m + 1 ].start )
                break;
            l = m + 1;
        }
        else
            r = m;
    }
    return (int64_t)m;
}

int64_t Index::seqIdForPositionOrRev( int64_t pos ) const // pack.h:1029-1034
{
    if( onReverse( pos ) )
        return seqIdForPosition( 2 * fwd_len - ( pos + 1 ) ) * 2 + 1;
    return seqIdForPosition( pos ) * 2;
}
int64_t Index::endOfSeqOrRev( int64_t id ) const // pack.h:1040-1045
{
    if( id % 2 == 1 )
        return ( 2 * fwd_len - ( contigs[ id / 2 ].start + 1 ) ) - 1;
    return contigs[ id / 2 ].start + contigs[ id / 2 ].length;
}
int64_t Index::startOfSeqOrRev( int64_t id ) const // pack.h:1047-1052
{
    if( id % 2 == 1 )
        return ( 2 * fwd_len - ( contigs[ id / 2 ].start + contigs[ id / 2 ].length + 1 ) ) + 1;
    return contigs[ id / 2 ].start;
}
bool Index::bridging( int64_t begin, int64_t size ) const // pack.h:1072-1087
{
    if( size <= 0 )
        return false;
    return onReverse( begin ) != onReverse( begin + size - 1 ) ||
           seqIdForPositionOrRev( begin ) != seqIdForPositionOrRev( begin + size - 1 );
}
void Index::extract( int64_t b, int64_t e, std::vector<uint8_t>& out ) const // pack.h:1147-1236 (holes ignored)
{
    if( b < 0 || b >= 2 * fwd_len || e < 0 || e > 2 * fwd_len )
        throw std::runtime_error( "(vExtractSubsection) range" );
    if( onReverse( b ) != onReverse( e - 1 ) )
        throw std::runtime_error( "(vExtractSubsection) Try to extract bridging sequence." );
    if( !( b <= e ) )
        throw std::runtime_error( "(vExtractSubsection) begin greater than end." );
    out.clear( );
    if( !onReverse( b ) )
        for( int64_t p = b; p < e; ++p )
            out.push_back( (uint8_t)nuc( p ) );
    else
        for( int64_t p = 2 * fwd_len - ( b + 1 ); p > 2 * fwd_len - ( e + 1 ); --p )
            out.push_back( (uint8_t)( 3 - nuc( p ) ) );
}

std::vector<Seed> extract_seeds( const Index& I, const Params& P, const std::vector<Segment>& segs, int64_t qlen,
                                 int64_t* pnInvPsi )
{
    std::vector<Seed> out;
    for( const Segment& s : segs ) // segment.h:316-349
    {
        if( (size_t)s.size < (size_t)P.min_seed_length )
            continue;
        if( s.sa.size > (int64_t)P.max_ambiguity && P.max_ambiguity != 0 )
            continue; // bSkip is hard-wired to true (segment.h:365)
        for( int64_t row = s.sa.start; row < s.sa.end( ); row++ ) // segment.h:89-113
        {
            if( pnInvPsi )
                *pnInvPsi += row & ( I.sa_intv - 1 ) ? 1 : 0; // lower bound only; exact count not needed here
            int64_t r = I.bwt_sa( row );
            const bool fw = r < I.ref_len / 2;
            if( !fw )
                r = I.ref_len - r - 1;
            Seed x;
            x.q = s.start, x.len = s.size + 1, x.r = r, x.amb = (unsigned)s.sa.size, x.fw = fw;
            // ExtractSeeds::setDeltaOfSeed, rectangular SoC (stripOfConsideration.h:42-54, 97-112)
            x.delta = x.r + ( qlen - x.q ) + ( qlen + 1 ) * I.seqIdForPosition( x.r );
            out.push_back( x );
        }
    }
    return out;
}

} // namespace oracle
