// Host side of the drop-in boundary, C++17, header only: the reference's Module interface for the path
// (libMS::Module<Ret, IS_VOLATILE, Args...>::execute, libs/ms/inc/ms/module/module.h:63-122), mirrored with the same
// module names, argument meaning and error behaviour (C++ exceptions), but over BATCHES of reads, because one GPU
// launch per read would waste the device.  Everything below only marshals std::vector buffers into
// the C ABI of libma_b200.so (include/ma_b200.h); no algorithm lives here and there is no CPU fallback.
//
//   reference (one read per call)                                          here (a batch per call)
//   BinarySeeding : Module<SegmentVector,false,SuffixArrayInterface,NucSeq>  BinarySeeding::execute(FMIndex, reads)
//     (binarySeeding.h:26, 571-584)                                          BinarySeeding::seed(FMIndex, reads)
//   StripOfConsideration : Module<SoCPriorityQueue,false,SegmentVector,NucSeq,Pack,FMIndex> (stripOfConsideration.h:164)
//   Harmonization : Module<ContainerVector<shared_ptr<Seeds>>,false,SoCPriorityQueue,NucSeq,FMIndex> (harmonization.h:34)
//                                                                           Harmonization::execute(FMIndex, reads)
//   NeedlemanWunsch : Module<ContainerVector<shared_ptr<Alignment>>,false,ContainerVector<shared_ptr<Seeds>>,NucSeq,Pack>
//     (needlemanWunsch.h:51, 111-134)                                       NeedlemanWunsch::execute(FMIndex, reads)
//   setUpCompGraph (export.cpp:72-128)                                      Aligner::align(reads)
#pragma once
#include "ma_b200.h"
#include <algorithm>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <limits>
#include <string>
#include <tuple>
#include <utility>
#include <vector>

namespace libMA_b200
{

typedef uint64_t nucSeqIndex;

// ParameterSetManager (parameter.h:1067-1201): presets by name, the selected set is what modules copy at construction
class ParameterSetManager
{
  public:
    ma_b200_params xParams;
    // "Detect Small Inversions" / "Z Drop Inversions" (parameter.h:640-647): read by the host-side SmallInversions only
    bool bSearchInversions = false;
    int iZDropInversion = 100;
    ParameterSetManager( )
    {
        ma_b200_params_preset( "default", &xParams );
    }
    void setSelected( const std::string& sName )
    {
        if( ma_b200_params_preset( sName.c_str( ), &xParams ) != MA_B200_OK )
            throw std::runtime_error( "unknown preset " + sName );
    }
};

// NucSeq (nucSeq.h:61-153): 1 byte per base, A=0 C=1 G=2 T=3 N=4
// Host buffer in page-locked memory (ma_b200_host_alloc) for the arrays that cross the C ABI in every batch; grows,
// never shrinks, resize() does not keep the contents.
template <typename T> class PinnedVector
{
    T* pData = nullptr;
    size_t uiSize = 0, uiCapacity = 0;

  public:
    PinnedVector( ) = default;
    PinnedVector( const PinnedVector& ) = delete;
    PinnedVector& operator=( const PinnedVector& ) = delete;
    PinnedVector( PinnedVector&& o ) noexcept : pData( o.pData ), uiSize( o.uiSize ), uiCapacity( o.uiCapacity )
    {
        o.pData = nullptr, o.uiSize = o.uiCapacity = 0;
    }
    PinnedVector& operator=( PinnedVector&& o ) noexcept
    {
        std::swap( pData, o.pData ), std::swap( uiSize, o.uiSize ), std::swap( uiCapacity, o.uiCapacity );
        return *this;
    }
    ~PinnedVector( )
    {
        ma_b200_host_free( pData );
    }
    void resize( size_t n )
    {
        if( n > uiCapacity )
        {
            ma_b200_host_free( pData );
            uiCapacity = n + n / 4 + 64;
            pData = (T*)ma_b200_host_alloc( (int64_t)( uiCapacity * sizeof( T ) ) );
            if( !pData )
            {
                uiCapacity = uiSize = 0;
                throw std::runtime_error( "ma_b200_host_alloc failed" );
            }
        }
        uiSize = n;
    }
    size_t size( ) const
    {
        return uiSize;
    }
    T* data( )
    {
        return pData;
    }
    const T* data( ) const
    {
        return pData;
    }
    T& operator[]( size_t i )
    {
        return pData[ i ];
    }
    const T& operator[]( size_t i ) const
    {
        return pData[ i ];
    }
};

class NucSeq
{
  public:
    std::vector<uint8_t> vSeq;
    std::vector<uint8_t> vQual; // ASCII qualities as read from FASTQ; empty = none (nucSeq.h WITH_QUALITY)
    std::string sName = "unknown";
    NucSeq( ) = default;
    explicit NucSeq( const std::string& sText )
    {
        for( char c : sText )
            vSeq.push_back( c == 'A' || c == 'a' ? 0 : c == 'C' || c == 'c' ? 1 : c == 'G' || c == 'g' ? 2
                                                   : c == 'T' || c == 't' ? 3 : 4 );
    }
    size_t length( ) const
    {
        return vSeq.size( );
    }
};

struct Segment // segment.h:31-115
{
    nucSeqIndex uiStart, uiSize; // size = length - 1
    int64_t iSaStart, iSaStartRevComp, iSaSize;
};
typedef std::vector<Segment> SegmentVector;

struct Seed // seed.h:34-43
{
    nucSeqIndex uiStart, uiSize, uiPosOnReference;
    unsigned int uiAmbiguity;
    bool bOnForwStrand;
    nucSeqIndex uiDelta;
};
struct Seeds
{
    std::vector<Seed> vContent;
    unsigned int index_of_strip = 0; // xStats.index_of_strip
};

enum MatchType // alignment.h:40-47
{
    seed,
    match,
    missmatch,
    insertion,
    deletion
};
struct Alignment // alignment.h:55-95
{
    std::vector<std::pair<MatchType, nucSeqIndex>> data;
    nucSeqIndex uiLength = 0, uiBeginOnRef = 0, uiEndOnRef = 0, uiBeginOnQuery = 0, uiEndOnQuery = 0;
    int64_t iScore = 0;
    unsigned int index_of_strip = 0;
    // set by MappingQuality / PairedReads (alignment.h:75-80, xStats.bFirst)
    double fMappingQuality = std::numeric_limits<double>::quiet_NaN( );
    bool bSecondary = false, bSupplementary = false, bFirst = false;
    int64_t score( ) const
    {
        return iScore;
    }
    // Alignment::append (alignment.cpp:11-98): run-length data, ends and the score with the capped affine gap cost
    void append( MatchType type, nucSeqIndex size, const ma_b200_params& P )
    {
        if( size == 0 )
            return;
        if( type == MatchType::seed || type == MatchType::match )
            iScore += (int64_t)P.match * (int64_t)size, uiEndOnRef += size, uiEndOnQuery += size;
        else if( type == MatchType::missmatch )
            iScore -= (int64_t)P.mismatch * (int64_t)size, uiEndOnRef += size, uiEndOnQuery += size;
        else
        {
            ( type == MatchType::insertion ? uiEndOnQuery : uiEndOnRef ) += size;
            auto cost = [ & ]( nucSeqIndex n ) {
                const nucSeqIndex c = (nucSeqIndex)P.extend * n + (nucSeqIndex)P.gap;
                return (int64_t)( c < (nucSeqIndex)P.sv_penalty ? c : (nucSeqIndex)P.sv_penalty );
            };
            if( !data.empty( ) && data.back( ).first == type )
            {
                size += data.back( ).second;
                uiLength -= data.back( ).second;
                iScore += cost( data.back( ).second );
                data.pop_back( );
            }
            iScore -= cost( size );
        }
        if( !data.empty( ) && data.back( ).first == type )
            data.back( ).second += size;
        else
            data.emplace_back( type, size );
        uiLength += size;
    }
};

// Pack's sequence descriptors (pack.h:39-176): name, start on the forward strand, length; read from <prefix>.ann
struct ContigTable
{
    std::vector<std::string> vNames;
    std::vector<int64_t> vStart, vLength;
    int64_t iForwardLength = 0;
    void vLoad( const std::string& sPrefix ) // Pack::vLoadCollection (pack.h:799-812), .ann part
    {
        std::ifstream ann( sPrefix + ".ann" );
        if( !ann )
            throw std::runtime_error( "File opening error: " + sPrefix + ".ann" );
        int64_t nSeq, seedv;
        ann >> iForwardLength >> nSeq >> seedv;
        if( !ann || iForwardLength < 0 || nSeq < 0 || nSeq > ( 1ll << 32 ) )
            throw std::runtime_error( "Corrupt file (header): " + sPrefix + ".ann" );
        vNames.clear( ), vStart.assign( nSeq, 0 ), vLength.assign( nSeq, 0 );
        std::string line;
        std::getline( ann, line );
        for( int64_t i = 0; i < nSeq; i++ )
        {
            std::getline( ann, line ); // "<gi> <name> <comment>"
            std::istringstream xL( line );
            std::string sGi, sName;
            xL >> sGi >> sName;
            vNames.push_back( sName );
            int64_t holes;
            ann >> vStart[ i ] >> vLength[ i ] >> holes;
            if( !ann || vStart[ i ] < 0 || vLength[ i ] < 0 || vStart[ i ] + vLength[ i ] > iForwardLength )
                throw std::runtime_error( "Corrupt file (contig " + std::to_string( i ) + "): " + sPrefix + ".ann" );
            std::getline( ann, line );
        }
    }
};

// One CUDA device: context + the replicated FMIndex / Pack (fMIndex.h, pack.h). Loads the reference's index files.
class FMIndex
{
    ma_b200_ctx* pCtx = nullptr;

  public:
    ContigTable xContigs; // Pack's sequence descriptors

    explicit FMIndex( int iDevice = 0 )
    {
        if( ma_b200_create( iDevice, &pCtx ) != MA_B200_OK )
            throw std::runtime_error( "ma_b200_create failed: no CUDA device (there is no CPU fallback)" );
    }
    FMIndex( const FMIndex& ) = delete;
    ~FMIndex( )
    {
        ma_b200_destroy( pCtx );
    }
    ma_b200_ctx* ctx( ) const
    {
        return pCtx;
    }
    void check( int rc ) const
    {
        if( rc != MA_B200_OK )
            throw std::runtime_error( std::string( "ma_b200: " ) + ma_b200_last_error( pCtx ) );
    }
    // FMIndex(prefix) + Pack(prefix): vLoadFMIndex (fMIndex.h:854-884), vLoadCollection (pack.h:799-812)
    void vLoad( const std::string& sPrefix )
    {
        auto slurp = []( const std::string& f ) {
            std::ifstream in( f, std::ios::binary | std::ios::ate );
            if( !in )
                throw std::runtime_error( "File opening error: " + f );
            std::vector<char> v( (size_t)in.tellg( ) );
            in.seekg( 0 );
            in.read( v.data( ), (std::streamsize)v.size( ) );
            return v;
        };
        // sizes derived from the headers are checked against the files before anything is copied (a truncated or
        // mismatched file must fail like the reference's loaders do, not read out of bounds)
        auto bad = [ & ]( const char* ext, const char* why ) {
            throw std::runtime_error( "Corrupt index file " + sPrefix + ext + ": " + why );
        };
        auto b = slurp( sPrefix + ".bwt" );
        if( b.size( ) < 40 || ( b.size( ) - 40 ) % 4 != 0 )
            bad( ".bwt", "shorter than its header or not a whole number of words" );
        int64_t primary, L2[ 5 ] = { 0, 0, 0, 0, 0 };
        memcpy( &primary, b.data( ), 8 );
        memcpy( &L2[ 1 ], b.data( ) + 8, 32 );
        const int64_t nWords = (int64_t)( b.size( ) - 40 ) / 4, refLen = L2[ 4 ];
        if( refLen <= 0 || L2[ 1 ] < 0 || L2[ 1 ] > L2[ 2 ] || L2[ 2 ] > L2[ 3 ] || L2[ 3 ] > L2[ 4 ] || primary < 0 || primary > refLen )
            bad( ".bwt", "inconsistent cumulative counts / primary" );
        if( nWords < ( refLen + 127 ) / 128 * 16 )
            bad( ".bwt", "fewer occurrence blocks than the reference length needs" );
        auto s = slurp( sPrefix + ".sa" );
        if( s.size( ) < 52 )
            bad( ".sa", "shorter than its header" );
        int32_t saIntv;
        memcpy( &saIntv, s.data( ) + 40, 4 );
        if( saIntv <= 0 || ( saIntv & ( saIntv - 1 ) ) != 0 )
            bad( ".sa", "sampling interval must be a positive power of two" );
        const int64_t nSa = ( refLen + saIntv ) / saIntv;
        if( (int64_t)s.size( ) < 52 + 8 * ( nSa - 1 ) )
            bad( ".sa", "fewer samples than the reference length needs" );
        std::vector<int64_t> sa( nSa );
        sa[ 0 ] = -1;
        memcpy( &sa[ 1 ], s.data( ) + 52, ( nSa - 1 ) * 8 );
        xContigs.vLoad( sPrefix );
        const std::vector<int64_t>&cs = xContigs.vStart, &cl = xContigs.vLength;
        const int64_t fwdLen = xContigs.iForwardLength, nSeq = (int64_t)cs.size( );
        auto p = slurp( sPrefix + ".pac" );
        if( refLen != 2 * fwdLen )
            bad( ".ann", "forward length does not match the BWT" );
        if( (int64_t)p.size( ) < ( fwdLen + 3 ) / 4 )
            bad( ".pac", "shorter than the forward strand" );
        check( ma_b200_index_upload( pCtx, (const uint32_t*)( b.data( ) + 40 ), nWords, L2, primary, refLen, sa.data( ),
                                     nSa, saIntv, (const uint8_t*)p.data( ), ( fwdLen + 3 ) / 4, fwdLen, cs.data( ),
                                     cl.data( ), (int32_t)nSeq ) );
        vPac.assign( p.begin( ), p.begin( ) + ( fwdLen + 3 ) / 4 );
    }
    // Pack::vExtract( begin, end ) (pack.h:1147-1236, 1440-1448; holes are not restored): [begin, end) of the forward
    // strand followed by its reverse complement; throws like the reference for ranges that bridge the strands
    std::vector<uint8_t> vExtract( int64_t iBegin, int64_t iEnd ) const
    {
        const int64_t iFwd = xContigs.iForwardLength, iTotal = 2 * iFwd;
        if( iBegin < 0 || iBegin >= iTotal || iEnd < 0 || iEnd > iTotal )
            throw std::runtime_error( "Pack (vExtractSubsection): out of range" );
        if( ( iBegin >= iFwd ) != ( iEnd - 1 >= iFwd ) )
            throw std::runtime_error( "(vExtractSubsection) Try to extract bridging sequence. This is impossible." );
        if( !( iBegin <= iEnd ) )
            throw std::runtime_error( "(vExtractSubsection) Try to extract with begin greater than end." );
        auto nuc = [ & ]( int64_t p ) { return (uint8_t)( ( vPac[ (size_t)( p >> 2 ) ] >> ( ( ~p & 3 ) << 1 ) ) & 3 ); };
        std::vector<uint8_t> v( (size_t)( iEnd - iBegin ) );
        if( iBegin < iFwd )
            for( int64_t p = iBegin; p < iEnd; p++ )
                v[ (size_t)( p - iBegin ) ] = nuc( p );
        else
            for( int64_t p = iBegin; p < iEnd; p++ )
                v[ (size_t)( p - iBegin ) ] = (uint8_t)( 3 - nuc( iTotal - 1 - p ) );
        return v;
    }

  private:
    std::vector<uint8_t> vPac; // host copy of the 2-bit forward strand (0.25 byte per base) for vExtract
};

namespace detail
{
inline void upload( FMIndex& rIdx, const ParameterSetManager& rP, const std::vector<NucSeq>& vReads )
{
    rIdx.check( ma_b200_set_params( rIdx.ctx( ), &rP.xParams ) );
    std::vector<uint8_t> vData;
    std::vector<int64_t> vOff( 1, 0 );
    for( auto& r : vReads )
    {
        vData.insert( vData.end( ), r.vSeq.begin( ), r.vSeq.end( ) );
        vOff.push_back( (int64_t)vData.size( ) );
    }
    vData.push_back( 0 );
    rIdx.check( ma_b200_align_upload( rIdx.ctx( ), (int64_t)vReads.size( ), vData.data( ), vOff.data( ) ) );
}
inline Seed toSeed( const ma_b200_seed& s )
{
    return Seed{ (nucSeqIndex)s.q, (nucSeqIndex)s.len, (nucSeqIndex)s.r, s.ambiguity, s.on_forward != 0,
                 (nucSeqIndex)s.delta };
}
} // namespace detail

class BinarySeeding
{
    const ParameterSetManager& rParams;

  public:
    explicit BinarySeeding( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    // BinarySeeding::execute for every read of the batch (binarySeeding.cpp:86-178)
    std::vector<SegmentVector> execute( FMIndex& rIdx, const std::vector<NucSeq>& vQueries, int iMaxSegments = 4096 )
    {
        detail::upload( rIdx, rParams, vQueries );
        ma_b200_align_stats st;
        rIdx.check( ma_b200_align_run( rIdx.ctx( ), MA_B200_STAGE_SEEDS, iMaxSegments, &st ) );
        std::vector<ma_b200_segment> vSeg( vQueries.size( ) * (size_t)iMaxSegments );
        std::vector<int32_t> vN( vQueries.size( ) );
        rIdx.check( ma_b200_align_download_segments( rIdx.ctx( ), vSeg.data( ), vN.data( ) ) );
        std::vector<SegmentVector> vRet( vQueries.size( ) );
        for( size_t i = 0; i < vQueries.size( ); i++ )
        {
            if( vN[ i ] > iMaxSegments )
                throw std::runtime_error( "BinarySeeding: more segments than iMaxSegments" );
            for( int j = 0; j < vN[ i ]; j++ )
            {
                const auto& s = vSeg[ i * (size_t)iMaxSegments + j ];
                vRet[ i ].push_back( Segment{ (nucSeqIndex)s.start, (nucSeqIndex)s.size, s.sa_start, s.sa_rev_start,
                                              s.sa_size } );
            }
        }
        return vRet;
    }
    // BinarySeeding::seed (binarySeeding.h:575-584): execute + extractSeeds, already batched in the reference
    std::vector<Seeds> seed( FMIndex& rIdx, const std::vector<NucSeq>& vQueries )
    {
        detail::upload( rIdx, rParams, vQueries );
        ma_b200_align_stats st;
        rIdx.check( ma_b200_align_run( rIdx.ctx( ), MA_B200_STAGE_SEEDS, 0, &st ) );
        std::vector<ma_b200_read_info> vInfo( vQueries.size( ) );
        std::vector<ma_b200_seed> vSeeds( (size_t)st.n_seeds + 1 );
        rIdx.check( ma_b200_align_download_info( rIdx.ctx( ), vInfo.data( ) ) );
        rIdx.check( ma_b200_align_download_seeds( rIdx.ctx( ), vSeeds.data( ), (int64_t)vSeeds.size( ) ) );
        std::vector<Seeds> vRet( vQueries.size( ) );
        for( size_t i = 0; i < vQueries.size( ); i++ )
            for( int j = 0; j < vInfo[ i ].n_seeds; j++ )
                vRet[ i ].vContent.push_back( detail::toSeed( vSeeds[ vInfo[ i ].seed_off + j ] ) );
        return vRet;
    }
};

// StripOfConsideration + Harmonization (the SoCPriorityQueue between them never leaves the device)
class Harmonization
{
    const ParameterSetManager& rParams;

  public:
    explicit Harmonization( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    std::vector<std::vector<Seeds>> execute( FMIndex& rIdx, const std::vector<NucSeq>& vQueries )
    {
        detail::upload( rIdx, rParams, vQueries );
        ma_b200_align_stats st;
        rIdx.check( ma_b200_align_run( rIdx.ctx( ), MA_B200_STAGE_SETS, 0, &st ) );
        std::vector<ma_b200_read_info> vInfo( vQueries.size( ) );
        std::vector<ma_b200_seed_set> vSets( (size_t)st.n_sets + 1 );
        std::vector<ma_b200_seed> vSeeds( (size_t)st.n_set_seeds + 1 );
        rIdx.check( ma_b200_align_download_info( rIdx.ctx( ), vInfo.data( ) ) );
        rIdx.check( ma_b200_align_download_sets( rIdx.ctx( ), vSets.data( ), (int64_t)vSets.size( ), vSeeds.data( ),
                                                 (int64_t)vSeeds.size( ) ) );
        std::vector<std::vector<Seeds>> vRet( vQueries.size( ) );
        for( size_t i = 0; i < vQueries.size( ); i++ )
            for( int k = 0; k < vInfo[ i ].n_sets; k++ )
            {
                const auto& h = vSets[ vInfo[ i ].set_off + k ];
                Seeds xS;
                xS.index_of_strip = h.soc_index;
                for( int j = 0; j < h.n; j++ )
                    xS.vContent.push_back( detail::toSeed( vSeeds[ h.seed_off + j ] ) );
                vRet[ i ].push_back( xS );
            }
        return vRet;
    }
};

class NeedlemanWunsch
{
    const ParameterSetManager& rParams;

  public:
    explicit NeedlemanWunsch( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    // per read: the alignments in the order of the reference's result vector (best first, needlemanWunsch.h:131-132)
    std::vector<std::vector<Alignment>> execute( FMIndex& rIdx, const std::vector<NucSeq>& vQueries,
                                                 ma_b200_align_stats* pStats = nullptr )
    {
        detail::upload( rIdx, rParams, vQueries );
        ma_b200_align_stats st;
        rIdx.check( ma_b200_align_run( rIdx.ctx( ), MA_B200_STAGE_ALIGN, 0, &st ) );
        if( pStats )
            *pStats = st;
        std::vector<ma_b200_read_info> vInfo( vQueries.size( ) );
        std::vector<ma_b200_alignment> vAln( (size_t)st.n_sets + 1 );
        std::vector<uint32_t> vRuns( (size_t)st.n_runs + 1 );
        rIdx.check( ma_b200_align_download( rIdx.ctx( ), vInfo.data( ), vAln.data( ), (int64_t)vAln.size( ),
                                            vRuns.data( ), (int64_t)vRuns.size( ) ) );
        std::vector<std::vector<Alignment>> vRet( vQueries.size( ) );
        for( size_t i = 0; i < vQueries.size( ); i++ )
        {
            vRet[ i ].resize( vInfo[ i ].n_sets );
            for( int k = 0; k < vInfo[ i ].n_sets; k++ )
            {
                const auto& a = vAln[ vInfo[ i ].set_off + k ];
                Alignment& x = vRet[ i ][ a.rank ];
                x.uiBeginOnRef = (nucSeqIndex)a.begin_ref, x.uiEndOnRef = (nucSeqIndex)a.end_ref;
                x.uiBeginOnQuery = (nucSeqIndex)a.begin_q, x.uiEndOnQuery = (nucSeqIndex)a.end_q;
                x.iScore = a.score, x.uiLength = (nucSeqIndex)a.length, x.index_of_strip = a.soc_index;
                for( int j = 0; j < a.n_runs; j++ )
                    x.data.emplace_back( (MatchType)( vRuns[ a.run_off + j ] & 7 ), vRuns[ a.run_off + j ] >> 3 );
            }
        }
        return vRet;
    }
};

namespace detail
{
// (into an existing object: its run vector keeps its capacity)
inline void fillAlignment( const ma_b200_alignment& a, const uint32_t* vRuns, Alignment& x )
{
    x.uiBeginOnRef = (nucSeqIndex)a.begin_ref, x.uiEndOnRef = (nucSeqIndex)a.end_ref;
    x.uiBeginOnQuery = (nucSeqIndex)a.begin_q, x.uiEndOnQuery = (nucSeqIndex)a.end_q;
    x.iScore = a.score, x.uiLength = (nucSeqIndex)a.length, x.index_of_strip = a.soc_index;
    x.fMappingQuality = a.mapq;
    x.bSecondary = ( a.flags & MA_B200_ALN_SECONDARY ) != 0, x.bSupplementary = ( a.flags & MA_B200_ALN_SUPPLEMENTARY ) != 0;
    x.bFirst = ( a.flags & MA_B200_ALN_FIRST_MATE ) != 0;
    x.data.clear( );
    x.data.reserve( (size_t)a.n_runs );
    for( int j = 0; j < a.n_runs; j++ )
        x.data.emplace_back( (MatchType)( vRuns[ a.run_off + j ] & 7 ), vRuns[ a.run_off + j ] >> 3 );
}
inline Alignment toAlignment( const ma_b200_alignment& a, const uint32_t* vRuns )
{
    Alignment x;
    fillAlignment( a, vRuns, x );
    return x;
}
// runs the path through MA_B200_STAGE_MAPQ and hands every alignment record to fVisit( read, record, runs )
template <typename F>
inline void runMapq( FMIndex& rIdx, ParameterSetManager xP, bool bPaired, const std::vector<NucSeq>& vQueries,
                     ma_b200_align_stats* pStats, F fVisit )
{
    xP.xParams.use_paired_reads = bPaired ? 1 : 0;
    upload( rIdx, xP, vQueries );
    ma_b200_align_stats st;
    rIdx.check( ma_b200_align_run( rIdx.ctx( ), MA_B200_STAGE_MAPQ, 0, &st ) );
    if( pStats )
        *pStats = st;
    std::vector<ma_b200_read_info> vInfo( vQueries.size( ) );
    std::vector<ma_b200_alignment> vAln( (size_t)st.n_sets + 1 );
    std::vector<uint32_t> vRuns( (size_t)st.n_runs + 1 );
    rIdx.check( ma_b200_align_download( rIdx.ctx( ), vInfo.data( ), vAln.data( ), (int64_t)vAln.size( ), vRuns.data( ),
                                        (int64_t)vRuns.size( ) ) );
    for( size_t i = 0; i < vQueries.size( ); i++ )
        for( int k = 0; k < vInfo[ i ].n_sets; k++ )
            fVisit( i, vAln[ vInfo[ i ].set_off + k ], vRuns.data( ) );
}
} // namespace detail

// The report of a batch as the C ABI delivers it: per-read info, alignment records, run words. records( i ) is what
// the reference's writer receives for read i (MappingQuality's vector) or, for paired presets, for pair i (PairedReads'
// vector over the mates 2i, 2i + 1).
struct RawReport
{
    PinnedVector<ma_b200_read_info> vInfo;
    PinnedVector<ma_b200_alignment> vAln;
    PinnedVector<uint32_t> vRuns;
    bool bPaired = false;

    size_t units( ) const
    {
        return bPaired ? vInfo.size( ) / 2 : vInfo.size( );
    }
    std::vector<Alignment> records( size_t i ) const
    {
        std::vector<Alignment> v;
        records( i, v );
        return v;
    }
    // into a caller-owned vector that is reused from unit to unit (no allocation once its elements have grown)
    void records( size_t i, std::vector<Alignment>& v ) const
    {
        const size_t uiFrom = bPaired ? 2 * i : i, uiTo = bPaired ? 2 * i + 2 : i + 1;
        size_t n = 0;
        for( size_t uiRead = uiFrom; uiRead < uiTo; uiRead++ )
            for( int k = 0; k < vInfo[ uiRead ].n_sets; k++ )
            {
                const auto& a = vAln[ vInfo[ uiRead ].set_off + k ];
                n = std::max( n, (size_t)( ( bPaired ? a.pair_rank : a.rank_mq ) + 1 ) );
            }
        if( v.size( ) < n )
            v.resize( n );
        else
            v.erase( v.begin( ) + n, v.end( ) );
        for( auto& x : v ) // a rank without a record (never produced by the device stages) reads as an empty alignment
            x.uiLength = 0, x.data.clear( );
        for( size_t uiRead = uiFrom; uiRead < uiTo; uiRead++ )
            for( int k = 0; k < vInfo[ uiRead ].n_sets; k++ )
            {
                const auto& a = vAln[ vInfo[ uiRead ].set_off + k ];
                const int iRank = bPaired ? a.pair_rank : a.rank_mq;
                if( iRank >= 0 )
                    detail::fillAlignment( a, vRuns.data( ), v[ (size_t)iRank ] );
            }
    }
};

// MappingQuality::execute for every read of the batch (mappingQuality.cpp:11-131): the reported alignments in the
// order of the reference's result vector, with bSecondary / bSupplementary / fMappingQuality set.
class MappingQuality
{
    const ParameterSetManager& rParams;

  public:
    explicit MappingQuality( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    std::vector<std::vector<Alignment>> execute( FMIndex& rIdx, const std::vector<NucSeq>& vQueries,
                                                 ma_b200_align_stats* pStats = nullptr )
    {
        std::vector<std::vector<Alignment>> vRet( vQueries.size( ) );
        detail::runMapq( rIdx, rParams, false, vQueries, pStats,
                         [ & ]( size_t i, const ma_b200_alignment& a, const uint32_t* vRuns ) {
                             if( a.rank_mq < 0 )
                                 return;
                             if( vRet[ i ].size( ) <= (size_t)a.rank_mq )
                                 vRet[ i ].resize( (size_t)a.rank_mq + 1 );
                             vRet[ i ][ a.rank_mq ] = detail::toAlignment( a, vRuns );
                         } );
        return vRet;
    }
};

// PairedReads::execute (pairedReads.cpp:15-121) for the mates vQueries[2k], vQueries[2k+1]: per pair the returned
// vector (the chosen alignment of each mate, or all alignments of the one mate that aligned).
class PairedReads
{
    const ParameterSetManager& rParams;

  public:
    explicit PairedReads( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    std::vector<std::vector<Alignment>> execute( FMIndex& rIdx, const std::vector<NucSeq>& vQueries,
                                                 ma_b200_align_stats* pStats = nullptr )
    {
        if( vQueries.size( ) % 2 )
            throw std::runtime_error( "PairedReads: the batch must hold the mates interleaved (2k, 2k+1)" );
        std::vector<std::vector<Alignment>> vRet( vQueries.size( ) / 2 );
        detail::runMapq( rIdx, rParams, true, vQueries, pStats,
                         [ & ]( size_t i, const ma_b200_alignment& a, const uint32_t* vRuns ) {
                             if( a.pair_rank < 0 )
                                 return;
                             auto& v = vRet[ i / 2 ];
                             if( v.size( ) <= (size_t)a.pair_rank )
                                 v.resize( (size_t)a.pair_rank + 1 );
                             v[ a.pair_rank ] = detail::toAlignment( a, vRuns );
                         } );
        return vRet;
    }
};

// SmallInversions::execute (smallInversions.h:22-221) for every read of a batch: vAlignments[ i ] is MappingQuality's
// vector of read i (the module sits between MappingQuality and the writer, export.cpp:109-112). Host glue as in the
// reference; the DP calls of the whole batch (kswcpp_dispatch, smallInversions.h:125-127) run as ONE ma_b200_ksw_batch.
class SmallInversions
{
    const ParameterSetManager& rParams;

    struct Candidate
    {
        size_t uiRead, uiAlignment;
        nucSeqIndex uiStartQ, uiEndQ, uiStartRRev;
        std::vector<uint8_t> vRef;
    };

    // The drop scan of smallInversions.h:54-115 in two steps. The alignment is cut at its seeds into windows — a seed and
    // everything up to the next seed (the leading window may start without one). Within a window the score of the path is
    // followed run by run from 0: `best` is the highest value so far (the position where it was LAST reached), and after
    // every run below it the drop is  best - score - extend * max(query, reference bases since best).  A window whose
    // largest drop reaches "Z Drop Inversions" and that is CLOSED by a seed is reported as the region between the end
    // of its own seed and the start of the closing one; the tail behind the last seed never is (the reference tests
    // the drop only when it meets a seed).
    struct RunStep // effect of one run of a given type: per-base query / reference advance, score per base and per run
    {
        int iQuery, iRef, iPerBase, iPerRun;
    };
    template <typename F> void forAllDropPos( const Alignment& a, F fDo ) const
    {
        const ma_b200_params& P = rParams.xParams;
        const RunStep aStep[ 5 ] = { { 1, 1, P.match, 0 }, // seed (scored like a match)
                                     { 1, 1, P.match, 0 }, // match
                                     { 1, 1, -P.mismatch, 0 }, // missmatch
                                     { 1, 0, -P.extend, -P.gap }, // insertion
                                     { 0, 1, -P.extend, -P.gap } }; // deletion
        const size_t n = a.data.size( );
        // positions in front of every run
        std::vector<nucSeqIndex> vQ( n + 1, a.uiBeginOnQuery ), vR( n + 1, a.uiBeginOnRef );
        for( size_t k = 0; k < n; k++ )
        {
            const RunStep& st = aStep[ (int)a.data[ k ].first ];
            vQ[ k + 1 ] = vQ[ k ] + (nucSeqIndex)st.iQuery * a.data[ k ].second;
            vR[ k + 1 ] = vR[ k ] + (nucSeqIndex)st.iRef * a.data[ k ].second;
        }
        size_t uiOpen = 0; // first run of the current window
        for( size_t uiClose = 0; uiClose < n; uiClose++ )
        {
            if( a.data[ uiClose ].first != MatchType::seed )
                continue;
            // window [uiOpen, uiClose), closed by the seed at uiClose (empty in front of a leading seed: drop 0)
            const bool bLeadingSeed = uiClose > uiOpen && a.data[ uiOpen ].first == MatchType::seed;
            long long iScore = 0, iBest = std::numeric_limits<int>::min( ), iDrop = 0;
            size_t uiBestAt = uiOpen; // index of the position (in front of run uiBestAt) of the best score
            for( size_t k = uiOpen; k < uiClose; k++ )
            {
                const RunStep& st = aStep[ (int)a.data[ k ].first ];
                iScore += (long long)st.iPerBase * (long long)a.data[ k ].second + st.iPerRun;
                if( iScore >= iBest )
                    iBest = iScore, uiBestAt = k + 1;
                else
                {
                    const nucSeqIndex uiDist = std::max( vQ[ k + 1 ] - vQ[ uiBestAt ], vR[ k + 1 ] - vR[ uiBestAt ] );
                    iDrop = std::max<long long>( iDrop, iBest - iScore - (long long)(int)uiDist * P.extend );
                }
            }
            if( iDrop >= rParams.iZDropInversion )
            {
                const size_t uiFrom = bLeadingSeed ? uiOpen + 1 : uiOpen; // region starts behind the window's own seed
                fDo( vQ[ uiFrom ], vR[ uiFrom ], vQ[ uiClose ], vR[ uiClose ] );
            }
            uiOpen = uiClose;
        }
    }

  public:
    explicit SmallInversions( const ParameterSetManager& rParameters ) : rParams( rParameters )
    {}
    std::vector<std::vector<Alignment>> execute( FMIndex& rIdx, const std::vector<std::vector<Alignment>>& vAlignments,
                                                 const std::vector<NucSeq>& vQueries )
    {
        const ma_b200_params& P = rParams.xParams;
        const nucSeqIndex uiTotal = 2 * (nucSeqIndex)rIdx.xContigs.iForwardLength;
        std::vector<Candidate> vCand;
        for( size_t i = 0; i < vAlignments.size( ); i++ )
            for( size_t k = 0; k < vAlignments[ i ].size( ); k++ )
                forAllDropPos( vAlignments[ i ][ k ],
                               [ & ]( nucSeqIndex uiStartQ, nucSeqIndex uiStartR, nucSeqIndex uiEndQ, nucSeqIndex uiEndR ) {
                                   // the window on the other strand (Pack::uiPositionToReverseStrand, pack.h:924-927)
                                   const nucSeqIndex uiStartRRev = uiTotal - ( uiEndR + 1 ),
                                                     uiEndRRev = uiTotal - ( uiStartR + 1 );
                                   vCand.push_back( Candidate{ i, k, uiStartQ, uiEndQ, uiStartRRev,
                                                               rIdx.vExtract( (int64_t)uiStartRRev, (int64_t)uiEndRRev ) } );
                               } );
        // tryInversionExtension (smallInversions.h:117-171): global-mode kswcpp call with the extension bandwidth
        std::vector<ma_b200_ksw_task> vTasks( vCand.size( ) );
        std::vector<uint8_t> vSeq;
        size_t uiCigarCap = 16;
        for( size_t c = 0; c < vCand.size( ); c++ )
        {
            const Candidate& x = vCand[ c ];
            const NucSeq& q = vQueries[ x.uiRead ];
            ma_b200_ksw_task& t = vTasks[ c ];
            t.qoff = (int64_t)vSeq.size( ), t.qlen = (int32_t)( (int)x.uiEndQ - (int)x.uiStartQ );
            vSeq.insert( vSeq.end( ), q.vSeq.begin( ) + x.uiStartQ, q.vSeq.begin( ) + x.uiEndQ );
            t.toff = (int64_t)vSeq.size( ), t.tlen = (int32_t)x.vRef.size( );
            vSeq.insert( vSeq.end( ), x.vRef.begin( ), x.vRef.end( ) );
            t.w = P.bandwidth_ext, t.zdrop = P.zdrop, t.flag = 0, t.tag = 0;
            uiCigarCap += (size_t)t.qlen + (size_t)t.tlen + 8;
        }
        std::vector<ma_b200_ksw_result> vRes( vCand.size( ) + 1 );
        std::vector<uint32_t> vCigar( uiCigarCap );
        int64_t iWords = 0;
        if( !vCand.empty( ) )
        {
            vSeq.push_back( 0 );
            rIdx.check( ma_b200_set_params( rIdx.ctx( ), &P ) );
            rIdx.check( ma_b200_ksw_batch( rIdx.ctx( ), (int64_t)vTasks.size( ), vTasks.data( ), vSeq.data( ),
                                           (int64_t)vSeq.size( ), vRes.data( ), vCigar.data( ), (int64_t)vCigar.size( ),
                                           &iWords ) );
        }
        std::vector<std::vector<Alignment>> vRet( vAlignments.size( ) );
        size_t c = 0;
        for( size_t i = 0; i < vAlignments.size( ); i++ )
            for( size_t k = 0; k < vAlignments[ i ].size( ); k++ )
            {
                vRet[ i ].push_back( vAlignments[ i ][ k ] );
                for( ; c < vCand.size( ) && vCand[ c ].uiRead == i && vCand[ c ].uiAlignment == k; c++ )
                {
                    const Candidate& x = vCand[ c ];
                    const NucSeq& q = vQueries[ i ];
                    Alignment xInv;
                    nucSeqIndex qPos = x.uiStartQ, rPos = 0;
                    for( int j = 0; j < vRes[ c ].n_cigar; j++ )
                    {
                        const uint32_t uiWord = vCigar[ (size_t)vRes[ c ].cigar_off + j ];
                        const uint32_t uiSymbol = uiWord & 0xf, uiAmount = uiWord >> 4;
                        if( uiSymbol == 0 )
                        {
                            for( uint32_t u = 0; u < uiAmount; u++ )
                                xInv.append( q.vSeq[ u + qPos ] == x.vRef[ u + rPos ] ? MatchType::match
                                                                                      : MatchType::missmatch, 1, P );
                            qPos += uiAmount, rPos += uiAmount;
                        }
                        else if( uiSymbol == 1 )
                            xInv.append( MatchType::insertion, uiAmount, P ), qPos += uiAmount;
                        else
                            xInv.append( MatchType::deletion, uiAmount, P ), rPos += uiAmount;
                    }
                    if( P.disable_heuristics || xInv.score( ) > (int64_t)P.harm_score_min * P.match )
                    {
                        xInv.uiBeginOnQuery += x.uiStartQ, xInv.uiEndOnQuery += x.uiStartQ;
                        xInv.uiBeginOnRef += x.uiStartRRev, xInv.uiEndOnRef += x.uiStartRRev;
                        xInv.bSupplementary = true;
                        xInv.index_of_strip = vAlignments[ i ][ k ].index_of_strip; // xStats of the parent
                        xInv.bFirst = vAlignments[ i ][ k ].bFirst;
                        xInv.fMappingQuality = 0;
                        vRet[ i ].push_back( xInv );
                    }
                }
            }
        return vRet;
    }
};

// PairedReads for graphs that run a host module between MappingQuality and PairedReads (SmallInversions adds records,
// export.cpp:176-184). No second implementation of the module: the records go through ma_b200_paired_reads_host, i.e. the
// routine of the device stage (ma_b200/csrc/mapq.cuh) compiled for the host; this class only marshals.
class PairedReadsHost
{
    const ma_b200_params& P;
    const int64_t iForwardLength;

    // the records the routine works on: what it reads of an alignment, and the run words for the seed count
    static void toRecords( const std::vector<Alignment>& v, std::vector<ma_b200_alignment>& vRec, std::vector<uint32_t>& vRuns )
    {
        for( size_t k = 0; k < v.size( ); k++ )
        {
            ma_b200_alignment r;
            memset( &r, 0, sizeof( r ) );
            r.begin_ref = (int64_t)v[ k ].uiBeginOnRef, r.end_ref = (int64_t)v[ k ].uiEndOnRef, r.score = v[ k ].iScore;
            r.begin_q = (int32_t)v[ k ].uiBeginOnQuery, r.end_q = (int32_t)v[ k ].uiEndOnQuery;
            r.length = (int32_t)v[ k ].uiLength, r.soc_index = v[ k ].index_of_strip;
            r.run_off = (int64_t)vRuns.size( ), r.n_runs = (int32_t)v[ k ].data.size( );
            for( const auto& d : v[ k ].data )
                vRuns.push_back( (uint32_t)d.second << 3 | (uint32_t)d.first );
            r.rank = r.rank_mq = (int32_t)k, r.pair_rank = -1, r.mapq = v[ k ].fMappingQuality;
            r.flags = ( v[ k ].bSecondary ? MA_B200_ALN_SECONDARY : 0 ) | ( v[ k ].bSupplementary ? MA_B200_ALN_SUPPLEMENTARY : 0 );
            vRec.push_back( r );
        }
    }

  public:
    PairedReadsHost( const ParameterSetManager& rParameters, const FMIndex& rIdx )
        : P( rParameters.xParams ), iForwardLength( rIdx.xContigs.iForwardLength )
    {}
    std::vector<Alignment> execute( const NucSeq& q1, const NucSeq& q2, std::vector<Alignment> v1,
                                    std::vector<Alignment> v2 ) const
    {
        std::vector<ma_b200_alignment> vRec1, vRec2;
        std::vector<uint32_t> vRuns;
        toRecords( v1, vRec1, vRuns ), toRecords( v2, vRec2, vRuns );
        vRuns.push_back( 0 );
        const int n = ma_b200_paired_reads_host( &P, 2 * iForwardLength, vRec1.data( ), (int32_t)vRec1.size( ),
                                                 (int64_t)q1.length( ), vRec2.data( ), (int32_t)vRec2.size( ),
                                                 (int64_t)q2.length( ), vRuns.data( ) );
        if( n < 0 )
            throw std::runtime_error( "PairedReads: no candidate pair for two aligned mates" );
        std::vector<Alignment> vRet( (size_t)n );
        auto collect = [ & ]( std::vector<Alignment>& v, const std::vector<ma_b200_alignment>& vRec, bool bFirstMate ) {
            for( size_t k = 0; k < v.size( ); k++ )
            {
                v[ k ].bFirst = bFirstMate; // pairedReads.cpp:21-25: every alignment learns which mate it belongs to
                if( vRec[ k ].pair_rank < 0 )
                    continue;
                v[ k ].bSecondary = ( vRec[ k ].flags & MA_B200_ALN_SECONDARY ) != 0;
                v[ k ].bSupplementary = ( vRec[ k ].flags & MA_B200_ALN_SUPPLEMENTARY ) != 0;
                v[ k ].fMappingQuality = vRec[ k ].mapq;
                vRet[ (size_t)vRec[ k ].pair_rank ] = v[ k ];
            }
        };
        collect( v1, vRec1, true ), collect( v2, vRec2, false );
        return vRet;
    }
};

// The batched graph: what setUpCompGraph / setUpCompGraphPaired (export.cpp:72-202) wire per thread, executed for a
// whole batch on one GPU.
class Aligner
{
    ParameterSetManager xParams;
    FMIndex xIndex;

  public:
    Aligner( const std::string& sIndexPrefix, const std::string& sPreset, int iDevice = 0 ) : xIndex( iDevice )
    {
        xParams.setSelected( sPreset );
        xIndex.vLoad( sIndexPrefix );
    }
    ParameterSetManager& params( )
    {
        return xParams;
    }
    const FMIndex& index( ) const
    {
        return xIndex;
    }
    // SmallInversions over the records of a batch (unpaired graphs; one DP batch on this aligner's device)
    std::vector<std::vector<Alignment>> inversions( const std::vector<std::vector<Alignment>>& vRecords,
                                                    const std::vector<NucSeq>& vReads )
    {
        return SmallInversions( xParams ).execute( xIndex, vRecords, vReads );
    }
    // NeedlemanWunsch results per read
    std::vector<std::vector<Alignment>> align( const std::vector<NucSeq>& vReads, ma_b200_align_stats* pStats = nullptr )
    {
        return NeedlemanWunsch( xParams ).execute( xIndex, vReads, pStats );
    }
    // what the writer receives: MappingQuality's vector per read, or PairedReads' vector per pair for paired presets
    std::vector<std::vector<Alignment>> report( const std::vector<NucSeq>& vReads,
                                                ma_b200_align_stats* pStats = nullptr )
    {
        if( xParams.bSearchInversions )
            return reportWithInversions( vReads );
        const RawReport xRaw = reportRaw( vReads, pStats );
        std::vector<std::vector<Alignment>> vRet( xRaw.units( ) );
        for( size_t i = 0; i < vRet.size( ); i++ )
            vRet[ i ] = xRaw.records( i );
        return vRet;
    }
    // "Detect Small Inversions": MappingQuality on the device, SmallInversions (host glue, DP on the device) per read,
    // then the writer's input: those vectors (export.cpp:109-112) or, for paired reads, PairedReads of the two mates'
    // vectors on the host (export.cpp:176-184)
    std::vector<std::vector<Alignment>> reportWithInversions( const std::vector<NucSeq>& vReads )
    {
        const bool bPaired = xParams.xParams.use_paired_reads != 0;
        if( bPaired && vReads.size( ) % 2 )
            throw std::runtime_error( "PairedReads: the batch must hold the mates interleaved (2k, 2k+1)" );
        ParameterSetManager xUnpaired = xParams;
        xUnpaired.xParams.use_paired_reads = 0;
        auto vInv = SmallInversions( xParams ).execute( xIndex, MappingQuality( xUnpaired ).execute( xIndex, vReads ), vReads );
        if( !bPaired )
            return vInv;
        std::vector<std::vector<Alignment>> vRet( vReads.size( ) / 2 );
        PairedReadsHost xPairing( xParams, xIndex );
        for( size_t p = 0; p < vRet.size( ); p++ )
            vRet[ p ] = xPairing.execute( vReads[ 2 * p ], vReads[ 2 * p + 1 ], vInv[ 2 * p ], vInv[ 2 * p + 1 ] );
        return vRet;
    }
    // the same result as the C ABI delivers it (record arrays of the whole batch); RawReport::records( i ) converts
    // one read (or pair) and is safe to call from several host threads at once
    RawReport reportRaw( const std::vector<NucSeq>& vReads, ma_b200_align_stats* pStats = nullptr )
    {
        RawReport xRaw;
        reportRaw( vReads, xRaw, pStats );
        return xRaw;
    }
    // into a caller-owned RawReport whose buffers are reused from batch to batch
    void reportRaw( const std::vector<NucSeq>& vReads, RawReport& xRaw, ma_b200_align_stats* pStats = nullptr )
    {
        fillSlab( vReads, vSlab, vOffsets );
        reportRaw( vReads.size( ), vSlab.data( ), vOffsets.data( ), xRaw, pStats );
    }
    // the reads of a batch as the C ABI takes them: concatenated base codes + offsets (uiParts-th part iPart of the
    // copy, so that several host threads can share it; offsets must have been filled by slabOffsets)
    static void slabOffsets( const std::vector<NucSeq>& vReads, PinnedVector<uint8_t>& rSlab,
                             PinnedVector<int64_t>& rOffsets )
    {
        rOffsets.resize( vReads.size( ) + 1 );
        size_t uiBytes = 0;
        for( size_t i = 0; i < vReads.size( ); i++ )
            rOffsets[ i ] = (int64_t)uiBytes, uiBytes += vReads[ i ].vSeq.size( );
        rOffsets[ vReads.size( ) ] = (int64_t)uiBytes;
        rSlab.resize( uiBytes + 1 );
        rSlab[ uiBytes ] = 0;
    }
    static void slabCopy( const std::vector<NucSeq>& vReads, PinnedVector<uint8_t>& rSlab,
                          const PinnedVector<int64_t>& rOffsets, size_t iPart = 0, size_t uiParts = 1 )
    {
        for( size_t i = vReads.size( ) * iPart / uiParts; i < vReads.size( ) * ( iPart + 1 ) / uiParts; i++ )
            if( !vReads[ i ].vSeq.empty( ) )
                memcpy( rSlab.data( ) + rOffsets[ i ], vReads[ i ].vSeq.data( ), vReads[ i ].vSeq.size( ) );
    }
    static void fillSlab( const std::vector<NucSeq>& vReads, PinnedVector<uint8_t>& rSlab, PinnedVector<int64_t>& rOffsets )
    {
        slabOffsets( vReads, rSlab, rOffsets );
        slabCopy( vReads, rSlab, rOffsets );
    }
    // with the slab prepared by the caller (maCMD_b200's reader threads do that while the device is busy)
    void reportRaw( size_t uiReads, const uint8_t* pSlab, const int64_t* pOffsets, RawReport& xRaw,
                    ma_b200_align_stats* pStats = nullptr )
    {
        xRaw.bPaired = xParams.xParams.use_paired_reads != 0;
        if( xRaw.bPaired && uiReads % 2 )
            throw std::runtime_error( "PairedReads: the batch must hold the mates interleaved (2k, 2k+1)" );
        xIndex.check( ma_b200_set_params( xIndex.ctx( ), &xParams.xParams ) );
        xIndex.check( ma_b200_align_upload( xIndex.ctx( ), (int64_t)uiReads, pSlab, pOffsets ) );
        ma_b200_align_stats st;
        xIndex.check( ma_b200_align_run( xIndex.ctx( ), MA_B200_STAGE_MAPQ, 0, &st ) );
        if( pStats )
            *pStats = st;
        xRaw.vInfo.resize( uiReads );
        xRaw.vAln.resize( (size_t)st.n_sets + 1 );
        xRaw.vRuns.resize( (size_t)st.n_runs + 1 );
        xIndex.check( ma_b200_align_download( xIndex.ctx( ), xRaw.vInfo.data( ), xRaw.vAln.data( ),
                                              (int64_t)xRaw.vAln.size( ), xRaw.vRuns.data( ),
                                              (int64_t)xRaw.vRuns.size( ) ) );
    }

  private:
    PinnedVector<uint8_t> vSlab; // the reads of the current batch as the C ABI takes them, kept between batches
    PinnedVector<int64_t> vOffsets;
};

} // namespace libMA_b200
