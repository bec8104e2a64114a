// ref_dump — harness around the UNMODIFIED reference (ITBE-Lab/ma) compiled into oracle/_ref/libma_ref.so.
// TEST INFRASTRUCTURE ONLY: generates golden vectors / parity dumps and times the reference's CPU path.
// Nothing in the product library (ma_b200/csrc) links or calls this.
//
// Drives exactly the reference's own modules, in the order of setUpCompGraph
// (/root/reference/libs/ma/src/util/export.cpp:72-128):
//   BinarySeeding::execute            libs/ma/src/module/binarySeeding.cpp:86-178
//   ExtractSeeds::execute             libs/ma/inc/ma/module/stripOfConsideration.h:138-157
//   StripOfConsiderationSeeds::execute libs/ma/src/module/stripOfConsideration.cpp:12-161
//   Harmonization::execute            libs/ma/src/module/harmonization.cpp:374-555
//   NeedlemanWunsch::execute          libs/ma/inc/ma/module/needlemanWunsch.h:111-134
// and interposes kswcpp_sse_xx (libs/kswcpp/src/kswcpp_sse_xx.cpp:77) to log every DP call.
//
// Sub-commands:
//   ref_dump index <genome.txt> <prefix>
//        genome.txt: lines ">name" followed by one line of ACGT. Writes <prefix>.{pac,ann,amb,bwt,sa}
//   ref_dump align <prefix> <reads.txt> <preset> <out.dump> [srand_base]
//        reads.txt: one read (ACGTN) per line. Dumps every stage. If srand_base >= 0 the harness calls
//        srand(srand_base + read_index) right before Harmonization::execute.
//   ref_dump ksw <pairs.txt> <out.dump>
//        pairs.txt: lines "w zdrop flag QUERY TARGET" (sequences as digits 0-4). DP only.
//   ref_dump bench <prefix> <reads.txt> <preset> <threads> [srand_base]
//        times the five modules over all reads with <threads> host threads; prints one JSON line.
//   ref_dump kswbench <pairs.txt> <threads> <repeat>
#include <sstream>
#include "ma/container/fMIndex.h"
#include "ma/container/pack.h"
#include "ma/module/binarySeeding.h"
#include "ma/module/fileReader.h"
#include "ma/module/fileWriter.h"
#include "ma/module/harmonization.h"
#include "ma/module/mappingQuality.h"
#include "ma/module/needlemanWunsch.h"
#include "ma/module/pairedReads.h"
#include "ma/module/smallInversions.h"
#include "ma/module/stripOfConsideration.h"
#include <atomic>
#include <chrono>
#include <dlfcn.h>
#include <fstream>
#include <map>
#include <cstring>
#include <thread>

using namespace libMA;
using namespace libMS;

// ---------------------------------------------------------------- dump container
struct Dump
{
    std::list<std::pair<std::string, std::vector<int64_t>>> vArrays; // list: references stay valid
    std::vector<int64_t>& arr( const std::string& sName )
    {
        for( auto& rP : vArrays )
            if( rP.first == sName )
                return rP.second;
        vArrays.emplace_back( sName, std::vector<int64_t>( ) );
        return vArrays.back( ).second;
    }
    void write( const std::string& sFile )
    {
        std::ofstream xOut( sFile, std::ios::binary );
        xOut << "MADUMP1\n";
        for( auto& rP : vArrays )
        {
            xOut << rP.first << " " << rP.second.size( ) << "\n";
            xOut.write( (const char*)rP.second.data( ), rP.second.size( ) * sizeof( int64_t ) );
        }
    }
};

// ---------------------------------------------------------------- ksw interposer
struct KswLog
{
    std::vector<int64_t> vCalls; // 16 fields per call
    std::vector<int64_t> vSeq; // query then target, one base per entry
    std::vector<int64_t> vCigar;
};
static thread_local KswLog* pKswLog = nullptr;
static std::atomic<int64_t> iKswCells( 0 );

typedef void ( *ksw_fn_t )( int, const uint8_t*, int, const uint8_t*, const KswCppParam<5>&, int, int, int,
                            kswcpp_extz_t*, AlignedMemoryManager& );
static const char* sKswSym = "_Z13kswcpp_sse_xxiPKhiS0_RK11KswCppParamILj5EEiiiP13kswcpp_extz_tR20AlignedMemoryManager";

void kswcpp_sse_xx( int qlen, const uint8_t* query, int tlen, const uint8_t* target, const KswCppParam<5>& xParam,
                    int w, int zdrop, int flag, kswcpp_extz_t* ez, AlignedMemoryManager& rxMemManager )
{
    static ksw_fn_t fReal = (ksw_fn_t)dlsym( RTLD_NEXT, sKswSym );
    if( !fReal )
    {
        std::cerr << "ref_dump: cannot resolve the reference's kswcpp_sse_xx" << std::endl;
        abort( );
    }
    fReal( qlen, query, tlen, target, xParam, w, zdrop, flag, ez, rxMemManager );
    if( pKswLog )
    {
        auto& c = pKswLog->vCalls;
        int64_t a[ 16 ] = { qlen,      tlen,     w,       zdrop,     flag,      (int64_t)ez->max, (int64_t)ez->zdropped,
                            ez->max_q, ez->max_t, ez->mqe, ez->mqe_t, ez->mte,   ez->mte_q,
                            ez->score, ez->n_cigar, ez->reach_end };
        c.insert( c.end( ), a, a + 16 );
        for( int i = 0; i < qlen; i++ )
            pKswLog->vSeq.push_back( query[ i ] );
        for( int i = 0; i < tlen; i++ )
            pKswLog->vSeq.push_back( target[ i ] );
        for( int i = 0; i < ez->n_cigar; i++ )
            pKswLog->vCigar.push_back( ez->cigar[ i ] );
    }
}

// ---------------------------------------------------------------- helpers
static std::vector<std::string> readLines( const std::string& sFile )
{
    std::ifstream xIn( sFile );
    if( !xIn )
    {
        std::cerr << "cannot open " << sFile << std::endl;
        exit( 2 );
    }
    std::vector<std::string> vRet;
    std::string sLine;
    while( std::getline( xIn, sLine ) )
        if( !sLine.empty( ) )
            vRet.push_back( sLine );
    return vRet;
}

static void selectPreset( ParameterSetManager& rP, std::string sPreset )
{
    for( auto& c : sPreset )
        c = std::tolower( c );
    rP.setSelected( sPreset );
    // MA_REF_MIN_GENOME_SIZE: "Minimum Genome Size for Heuristics" (parameter.h:873, default 10 M): 0 switches the
    // heuristics for large genomes (seeding drop-off, SoC minimal length) on for the small golden genome
    if( const char* pMin = getenv( "MA_REF_MIN_GENOME_SIZE" ) )
        rP.getSelected( )->xGenomeSizeDisable->set( atoi( pMin ) );
    // MA_REF_SET="Name=value;Name=value": parameters of the selected presetting by their reference names, set the way
    // the reference's CLI does it (cmdMa.cpp:347-358: byName( ... )->setByText( ... ))
    if( const char* pSet = getenv( "MA_REF_SET" ) )
    {
        std::stringstream xS( pSet );
        std::string sItem;
        while( std::getline( xS, sItem, ';' ) )
        {
            const size_t uiEq = sItem.find( '=' );
            if( uiEq == std::string::npos )
                continue;
            auto pParam = rP.getSelected( )->byName( sItem.substr( 0, uiEq ) );
            const std::string sValue = sItem.substr( uiEq + 1 );
            // flags are set directly: the reference's text conversion returns true for "false" as well
            // (libs/ms/src/util/parameter.cpp:40-43), so that its CLI can only switch flags on
            if( auto pFlag = std::dynamic_pointer_cast<AlignerParameter<bool>>( pParam ) )
                pFlag->set( sValue == "true" );
            else
                pParam->setByText( sValue );
        }
    }
}

static int cmdIndex( int argc, char** argv )
{
    auto vLines = readLines( argv[ 2 ] );
    auto pPack = std::make_shared<Pack>( );
    for( size_t i = 0; i + 1 < vLines.size( ); i += 2 )
    {
        std::string sName = vLines[ i ].substr( 1 );
        NucSeq xSeq( vLines[ i + 1 ] );
        pPack->vAppendSequence( sName, "synthetic", xSeq );
    }
    pPack->vStoreCollection( argv[ 3 ] );
    auto pFM = std::make_shared<FMIndex>( pPack );
    pFM->vStoreFMIndex( argv[ 3 ] );
    return 0;
}

struct Stages
{
    std::shared_ptr<SegmentVector> pSegments;
    std::shared_ptr<Seeds> pSeeds;
    std::shared_ptr<SoCPriorityQueue> pSoCs;
    std::shared_ptr<ContainerVector<std::shared_ptr<Seeds>>> pHarm;
    std::shared_ptr<ContainerVector<std::shared_ptr<Alignment>>> pAlignments;
};

struct Modules
{
    BinarySeeding xSeeding;
    StripOfConsideration xSoC;
    Harmonization xHarm;
    NeedlemanWunsch xNW;
    MappingQuality xMQ;
    PairedReads xPR;
    Modules( const ParameterSetManager& rP )
        : xSeeding( rP ), xSoC( rP ), xHarm( rP ), xNW( rP ), xMQ( rP ), xPR( rP )
    {}
};

static int cmdAlign( int argc, char** argv )
{
    std::string sPrefix = argv[ 2 ];
    auto vReads = readLines( argv[ 3 ] );
    ParameterSetManager xP;
    selectPreset( xP, argv[ 4 ] );
    int64_t iSrandBase = argc > 6 ? atoll( argv[ 6 ] ) : -1;
    auto pPack = std::make_shared<Pack>( sPrefix );
    auto pFM = std::make_shared<FMIndex>( sPrefix );
    Modules xM( xP );

    Dump xD;
    auto& seg_off = xD.arr( "seg_off" );
    auto& seg = xD.arr( "seg" );
    auto& seed_off = xD.arr( "seed_off" );
    auto& seed = xD.arr( "seed" );
    auto& soc_off = xD.arr( "soc_off" );
    auto& soc = xD.arr( "soc" );
    auto& socseed_off = xD.arr( "socseed_off" );
    auto& socseed = xD.arr( "socseed" );
    auto& harm_off = xD.arr( "harm_off" );
    auto& harmseed_off = xD.arr( "harmseed_off" );
    auto& harmseed = xD.arr( "harmseed" );
    auto& aln_off = xD.arr( "aln_off" );
    auto& aln = xD.arr( "aln" );
    auto& alndata_off = xD.arr( "alndata_off" );
    auto& alndata = xD.arr( "alndata" );
    auto& ksw_off = xD.arr( "ksw_off" );
    // MappingQuality per read: rows {index in the NeedlemanWunsch result, secondary | supplementary << 1, bits of the
    // double mapping quality}, in the order of the returned vector. PairedReads per pair of consecutive reads
    // (2k, 2k+1): rows {mate (0 / 1), index in that mate's NeedlemanWunsch result, flags, mapping quality bits}.
    auto& mq_off = xD.arr( "mq_off" );
    auto& mq = xD.arr( "mq" );
    auto& pr_off = xD.arr( "pr_off" );
    auto& pr = xD.arr( "pr" );
    mq_off.push_back( 0 );
    pr_off.push_back( 0 );
    std::shared_ptr<NucSeq> pPrevQ;
    std::shared_ptr<ContainerVector<std::shared_ptr<Alignment>>> pPrevMQ;
    std::map<const Alignment*, int64_t> xPrevIdx;
    KswLog xLog;
    seg_off.push_back( 0 );
    seed_off.push_back( 0 );
    soc_off.push_back( 0 );
    socseed_off.push_back( 0 );
    harm_off.push_back( 0 );
    harmseed_off.push_back( 0 );
    aln_off.push_back( 0 );
    alndata_off.push_back( 0 );
    ksw_off.push_back( 0 );

    for( size_t uiRead = 0; uiRead < vReads.size( ); uiRead++ )
    {
        auto pQ = std::make_shared<NucSeq>( vReads[ uiRead ] );
        pQ->sName = "r" + std::to_string( uiRead );
        auto pSeg = xM.xSeeding.execute( pFM, pQ );
        for( const Segment& rS : *pSeg )
        {
            int64_t a[ 5 ] = { (int64_t)rS.start( ), (int64_t)rS.size( ), rS.saInterval( ).start( ),
                               rS.saInterval( ).startRevComp( ), rS.saInterval( ).size( ) };
            seg.insert( seg.end( ), a, a + 5 );
        }
        seg_off.push_back( seg.size( ) / 5 );

        auto pSeeds = xM.xSoC.xExtractHelper.execute( pSeg, pFM, pQ, pPack );
        for( const Seed& rS : *pSeeds )
        {
            int64_t a[ 6 ] = { (int64_t)rS.start( ),  (int64_t)rS.size( ),         (int64_t)rS.start_ref( ),
                               (int64_t)rS.uiAmbiguity, (int64_t)rS.bOnForwStrand, (int64_t)rS.uiDelta };
            seed.insert( seed.end( ), a, a + 6 );
        }
        seed_off.push_back( seed.size( ) / 6 );

        auto pSoCs = xM.xSoC.xHelper.execute( pSeeds, pQ, pPack );
        {
            // pop a deep copy of the queue to record the complete pop order without disturbing the original
            auto pSeedCopy = std::make_shared<Seeds>( pSoCs->pSeeds );
            SoCPriorityQueue xCopy( pSeedCopy );
            for( auto& rT : pSoCs->vMaxima )
                xCopy.vMaxima.emplace_back( std::get<0>( rT ),
                                            pSeedCopy->begin( ) + ( std::get<1>( rT ) - pSoCs->pSeeds->begin( ) ),
                                            pSeedCopy->begin( ) + ( std::get<2>( rT ) - pSoCs->pSeeds->begin( ) ) );
            while( !xCopy.empty( ) )
            {
                SoCOrder xO = std::get<0>( xCopy.vMaxima.front( ) );
                auto pPop = xCopy.pop( );
                int64_t a[ 4 ] = { (int64_t)xO.uiAccumulativeLength, (int64_t)xO.uiSeedAmbiguity,
                                   (int64_t)xO.uiSeedAmount, (int64_t)pPop->xStats.index_of_strip };
                soc.insert( soc.end( ), a, a + 4 );
                for( const Seed& rS : *pPop )
                {
                    int64_t b[ 4 ] = { (int64_t)rS.start( ), (int64_t)rS.size( ), (int64_t)rS.start_ref( ),
                                       (int64_t)rS.bOnForwStrand };
                    socseed.insert( socseed.end( ), b, b + 4 );
                }
                socseed_off.push_back( socseed.size( ) / 4 );
            }
        }
        soc_off.push_back( soc.size( ) / 4 );

        if( iSrandBase >= 0 )
            srand( (unsigned int)( iSrandBase + uiRead ) );
        auto pHarm = xM.xHarm.execute( pSoCs, pQ, pFM );
        for( auto pSet : *pHarm )
        {
            for( const Seed& rS : *pSet )
            {
                int64_t b[ 5 ] = { (int64_t)rS.start( ), (int64_t)rS.size( ), (int64_t)rS.start_ref( ),
                                   (int64_t)rS.bOnForwStrand, (int64_t)pSet->xStats.index_of_strip };
                harmseed.insert( harmseed.end( ), b, b + 5 );
            }
            harmseed_off.push_back( harmseed.size( ) / 5 );
        }
        harm_off.push_back( harmseed_off.size( ) - 1 );

        pKswLog = &xLog;
        auto pAln = xM.xNW.execute( pHarm, pQ, pPack );
        pKswLog = nullptr;
        ksw_off.push_back( xLog.vCalls.size( ) / 16 );
        for( auto pA : *pAln )
        {
            int64_t a[ 8 ] = { (int64_t)pA->uiBeginOnQuery, (int64_t)pA->uiEndOnQuery,
                               (int64_t)pA->uiBeginOnRef,   (int64_t)pA->uiEndOnRef,
                               pA->iScore,                  (int64_t)pA->xStats.index_of_strip,
                               (int64_t)pA->uiLength,       (int64_t)pA->data.size( ) };
            aln.insert( aln.end( ), a, a + 8 );
            for( auto& rP : pA->data )
            {
                alndata.push_back( (int64_t)rP.first );
                alndata.push_back( (int64_t)rP.second );
            }
            alndata_off.push_back( alndata.size( ) / 2 );
        }
        aln_off.push_back( aln.size( ) / 8 );

        // ---- MappingQuality (mappingQuality.cpp:11-131) on the NeedlemanWunsch result of the read
        std::map<const Alignment*, int64_t> xIdx;
        for( size_t i = 0; i < pAln->size( ); i++ )
            xIdx[ ( *pAln )[ i ].get( ) ] = (int64_t)i;
        auto fBits = []( double d ) {
            int64_t i;
            memcpy( &i, &d, 8 );
            return i;
        };
        auto pMQ = xM.xMQ.execute( pQ, pAln );
        for( auto pA : *pMQ )
        {
            int64_t a[ 3 ] = { xIdx.at( pA.get( ) ), (int64_t)pA->bSecondary | ( (int64_t)pA->bSupplementary << 1 ),
                               fBits( pA->fMappingQuality ) };
            mq.insert( mq.end( ), a, a + 3 );
        }
        mq_off.push_back( mq.size( ) / 3 );
        // ---- PairedReads (pairedReads.cpp:15-121) on every pair of consecutive reads
        if( uiRead % 2 == 0 )
            pPrevQ = pQ, pPrevMQ = pMQ, xPrevIdx = xIdx;
        else
        {
            auto pPR = xM.xPR.execute( pPrevQ, pQ, pPrevMQ, pMQ, pPack );
            for( auto pA : *pPR )
            {
                const bool bSecondMate = xIdx.count( pA.get( ) ) != 0;
                int64_t a[ 4 ] = { (int64_t)bSecondMate, bSecondMate ? xIdx.at( pA.get( ) ) : xPrevIdx.at( pA.get( ) ),
                                   (int64_t)pA->bSecondary | ( (int64_t)pA->bSupplementary << 1 ),
                                   fBits( pA->fMappingQuality ) };
                pr.insert( pr.end( ), a, a + 4 );
            }
            pr_off.push_back( pr.size( ) / 4 );
        }
    }
    xD.arr( "ksw_calls" ) = xLog.vCalls;
    xD.arr( "ksw_seq" ) = xLog.vSeq;
    xD.arr( "ksw_cigar" ) = xLog.vCigar;
    xD.write( argv[ 5 ] );
    return 0;
}

// ref_dump sam <index prefix> <reads.txt> <preset> <out.sam> [srand_base [inv]]
// The reference's own FileWriter / PairedFileWriter (fileWriter.cpp:11-156, 158-372) behind the path, reads named r<i>;
// for presets with "Use Paired Reads" the reads 2k, 2k+1 are mates.
// reads of a FASTA / FASTQ file through the reference's own FileReader (fileReader.cpp:37-203); plain sequence lines
// (the fixture format of this repo) are named r<i>
static std::vector<std::shared_ptr<NucSeq>> readQueries( const ParameterSetManager& rP, const std::string& sFile )
{
    std::vector<std::shared_ptr<NucSeq>> vRet;
    std::ifstream xProbe( sFile );
    const int c = xProbe.peek( );
    if( c == '>' || c == '@' )
    {
        FileReader xReader( rP );
        auto pStream = std::make_shared<FileStreamFromPath>( sFile );
        while( auto pQ = xReader.execute( pStream ) )
            vRet.push_back( pQ );
        return vRet;
    }
    auto vLines = readLines( sFile );
    for( size_t i = 0; i < vLines.size( ); i++ )
    {
        vRet.push_back( std::make_shared<NucSeq>( vLines[ i ] ) );
        vRet.back( )->sName = "r" + std::to_string( i );
    }
    return vRet;
}

// ref_dump reads <file>: name, sequence and quality of every read as the reference's FileReader delivers them
static int cmdReads( int argc, char** argv )
{
    ParameterSetManager xP;
    for( auto pQ : readQueries( xP, argv[ 2 ] ) )
        std::cout << pQ->sName << "\t" << pQ->toString( ) << "\t" << pQ->toQualString( ) << "\n";
    return 0;
}

static int cmdSam( int argc, char** argv )
{
    std::string sPrefix = argv[ 2 ];
    ParameterSetManager xP;
    selectPreset( xP, argv[ 4 ] );
    auto vReads = readQueries( xP, argv[ 3 ] );
    int64_t iSrandBase = argc > 6 ? atoll( argv[ 6 ] ) : -1;
    // "inv": "Detect Small Inversions" on, SmallInversions between MappingQuality and the writer (export.cpp:109-112)
    // "inv" or "inv=<Z Drop Inversions>"
    const bool bInversions = argc > 7 && std::string( argv[ 7 ] ).substr( 0, 3 ) == "inv";
    if( bInversions && std::string( argv[ 7 ] ).size( ) > 4 )
        xP.getSelected( )->xZDropInversion->set( atoi( argv[ 7 ] + 4 ) );
    auto pPack = std::make_shared<Pack>( sPrefix );
    auto pFM = std::make_shared<FMIndex>( sPrefix );
    Modules xM( xP );
    SmallInversions xInv( xP );
    const bool bPaired = xP.getSelected( )->xUsePairedReads->get( );
    std::shared_ptr<FileWriter> pW;
    std::shared_ptr<PairedFileWriter> pPW;
    if( bPaired )
        pPW = std::make_shared<PairedFileWriter>( xP, std::string( argv[ 5 ] ), pPack );
    else
        pW = std::make_shared<FileWriter>( xP, std::string( argv[ 5 ] ), pPack );
    std::shared_ptr<NucSeq> pPrevQ;
    std::shared_ptr<ContainerVector<std::shared_ptr<Alignment>>> pPrevMQ;
    for( size_t uiRead = 0; uiRead < vReads.size( ); uiRead++ )
    {
        auto pQ = vReads[ uiRead ];
        auto pSeg = xM.xSeeding.execute( pFM, pQ );
        auto pSeeds = xM.xSoC.xExtractHelper.execute( pSeg, pFM, pQ, pPack );
        auto pSoCs = xM.xSoC.xHelper.execute( pSeeds, pQ, pPack );
        if( iSrandBase >= 0 )
            srand( (unsigned int)( iSrandBase + uiRead ) );
        auto pHarm = xM.xHarm.execute( pSoCs, pQ, pFM );
        auto pAln = xM.xNW.execute( pHarm, pQ, pPack );
        auto pMQ = xM.xMQ.execute( pQ, pAln );
        if( bInversions )
            pMQ = xInv.execute( pMQ, pQ, pPack );
        if( !bPaired )
            pW->execute( pQ, pMQ, pPack );
        else if( uiRead % 2 == 0 )
            pPrevQ = pQ, pPrevMQ = pMQ;
        else
            pPW->execute( pPrevQ, pQ, xM.xPR.execute( pPrevQ, pQ, pPrevMQ, pMQ, pPack ), pPack );
    }
    return 0;
}

struct KswPair
{
    int w, zdrop, flag;
    std::vector<uint8_t> q, t;
};
static std::vector<KswPair> readPairs( const std::string& sFile )
{
    std::vector<KswPair> vRet;
    std::ifstream xIn( sFile );
    KswPair xP;
    std::string sQ, sT;
    while( xIn >> xP.w >> xP.zdrop >> xP.flag >> sQ >> sT )
    {
        xP.q.clear( );
        xP.t.clear( );
        for( char c : sQ )
            if( c != '-' )
                xP.q.push_back( c - '0' );
        for( char c : sT )
            if( c != '-' )
                xP.t.push_back( c - '0' );
        vRet.push_back( xP );
    }
    return vRet;
}

static int cmdKsw( int argc, char** argv )
{
    auto vPairs = readPairs( argv[ 2 ] );
    // optional: match mismatch gap extend gap2 extend2 (KswCppParam<5>, kswcpp.h:44-129); default = pGlobalParams
    int aS[ 6 ] = { 2, 4, 4, 2, 24, 1 };
    for( int i = 0; i < 6 && argc > 4 + i; i++ )
        aS[ i ] = atoi( argv[ 4 + i ] );
    KswCppParam<5> xParam( aS[ 0 ], aS[ 1 ], aS[ 2 ], aS[ 3 ], aS[ 4 ], aS[ 5 ] );
    KswLog xLog;
    AlignedMemoryManager xMem;
    pKswLog = &xLog;
    for( auto& rP : vPairs )
    {
        Wrapper_ksw_extz_t ez;
        kswcpp_dispatch( (int)rP.q.size( ), rP.q.data( ), (int)rP.t.size( ), rP.t.data( ), xParam, rP.w, rP.zdrop,
                         rP.flag, ez.ez, xMem );
    }
    pKswLog = nullptr;
    Dump xD;
    xD.arr( "ksw_calls" ) = xLog.vCalls;
    xD.arr( "ksw_seq" ) = xLog.vSeq;
    xD.arr( "ksw_cigar" ) = xLog.vCigar;
    xD.write( argv[ 3 ] );
    return 0;
}

static int cmdBench( int argc, char** argv )
{
    std::string sPrefix = argv[ 2 ];
    auto vReads = readLines( argv[ 3 ] );
    ParameterSetManager xP;
    selectPreset( xP, argv[ 4 ] );
    int iThreads = atoi( argv[ 5 ] );
    int64_t iSrandBase = argc > 6 ? atoll( argv[ 6 ] ) : -1;
    auto pPack = std::make_shared<Pack>( sPrefix );
    auto pFM = std::make_shared<FMIndex>( sPrefix );
    Modules xM( xP );
    std::vector<std::shared_ptr<NucSeq>> vQ;
    for( auto& s : vReads )
        vQ.push_back( std::make_shared<NucSeq>( s ) );
    std::atomic<size_t> uiNext( 0 );
    std::atomic<size_t> uiAligned( 0 );
    std::atomic<int64_t> iNs[ 6 ];
    for( auto& x : iNs )
        x = 0;
    const bool bPaired = xP.getSelected( )->xUsePairedReads->get( );
    auto tStart = std::chrono::steady_clock::now( );
    std::vector<std::thread> vT;
    for( int t = 0; t < iThreads; t++ )
        vT.emplace_back( [ & ]( ) {
            int64_t aNs[ 6 ] = { 0, 0, 0, 0, 0, 0 };
            size_t uiLocalAligned = 0;
            while( true )
            {
                size_t i = uiNext.fetch_add( 16 ); // even: mates (2k, 2k+1) stay in one chunk
                if( i >= vQ.size( ) )
                    break;
                std::shared_ptr<ContainerVector<std::shared_ptr<Alignment>>> pPrevMQ;
                for( size_t j = i; j < std::min( i + 16, vQ.size( ) ); j++ )
                {
                    auto t0 = std::chrono::steady_clock::now( );
                    auto pSeg = xM.xSeeding.execute( pFM, vQ[ j ] );
                    auto t1 = std::chrono::steady_clock::now( );
                    auto pSeeds = xM.xSoC.xExtractHelper.execute( pSeg, pFM, vQ[ j ], pPack );
                    auto t2 = std::chrono::steady_clock::now( );
                    auto pSoCs = xM.xSoC.xHelper.execute( pSeeds, vQ[ j ], pPack );
                    auto t3 = std::chrono::steady_clock::now( );
                    if( iSrandBase >= 0 && iThreads == 1 )
                        srand( (unsigned int)( iSrandBase + j ) );
                    auto pHarm = xM.xHarm.execute( pSoCs, vQ[ j ], pFM );
                    auto t4 = std::chrono::steady_clock::now( );
                    auto pAln = xM.xNW.execute( pHarm, vQ[ j ], pPack );
                    auto t5 = std::chrono::steady_clock::now( );
                    if( !pAln->empty( ) && ( *pAln )[ 0 ]->uiLength > 0 )
                        uiLocalAligned++;
                    auto pMQ = xM.xMQ.execute( vQ[ j ], pAln );
                    if( bPaired )
                    {
                        if( j % 2 == 0 )
                            pPrevMQ = pMQ;
                        else
                            xM.xPR.execute( vQ[ j - 1 ], vQ[ j ], pPrevMQ, pMQ, pPack );
                    }
                    auto t6 = std::chrono::steady_clock::now( );
                    aNs[ 0 ] += ( t1 - t0 ).count( );
                    aNs[ 1 ] += ( t2 - t1 ).count( );
                    aNs[ 2 ] += ( t3 - t2 ).count( );
                    aNs[ 3 ] += ( t4 - t3 ).count( );
                    aNs[ 4 ] += ( t5 - t4 ).count( );
                    aNs[ 5 ] += ( t6 - t5 ).count( );
                }
            }
            for( int k = 0; k < 6; k++ )
                iNs[ k ] += aNs[ k ];
            uiAligned += uiLocalAligned;
        } );
    for( auto& t : vT )
        t.join( );
    double fSec = std::chrono::duration<double>( std::chrono::steady_clock::now( ) - tStart ).count( );
    printf( "{\"reads\": %zu, \"aligned\": %zu, \"threads\": %d, \"seconds\": %.6f, \"reads_per_s\": %.1f, "
            "\"stage_cpu_s\": {\"seeding\": %.4f, \"extract\": %.4f, \"soc\": %.4f, \"harmonization\": %.4f, \"dp\": "
            "%.4f, \"mapq_pairing\": %.4f}}\n",
            vQ.size( ), (size_t)uiAligned, iThreads, fSec, vQ.size( ) / fSec, iNs[ 0 ] * 1e-9, iNs[ 1 ] * 1e-9,
            iNs[ 2 ] * 1e-9, iNs[ 3 ] * 1e-9, iNs[ 4 ] * 1e-9, iNs[ 5 ] * 1e-9 );
    return 0;
}

static int cmdKswBench( int argc, char** argv )
{
    auto vPairs = readPairs( argv[ 2 ] );
    int iThreads = atoi( argv[ 3 ] );
    int iRepeat = atoi( argv[ 4 ] );
    KswCppParam<5> xParam( 2, 4, 4, 2, 24, 1 );
    std::atomic<size_t> uiNext( 0 );
    size_t uiTotal = vPairs.size( ) * (size_t)iRepeat;
    auto tStart = std::chrono::steady_clock::now( );
    std::vector<std::thread> vT;
    for( int t = 0; t < iThreads; t++ )
        vT.emplace_back( [ & ]( ) {
            AlignedMemoryManager xMem;
            while( true )
            {
                size_t i = uiNext.fetch_add( 4 );
                if( i >= uiTotal )
                    break;
                for( size_t j = i; j < std::min( i + 4, uiTotal ); j++ )
                {
                    auto& rP = vPairs[ j % vPairs.size( ) ];
                    Wrapper_ksw_extz_t ez;
                    kswcpp_dispatch( (int)rP.q.size( ), rP.q.data( ), (int)rP.t.size( ), rP.t.data( ), xParam, rP.w,
                                     rP.zdrop, rP.flag, ez.ez, xMem );
                }
            }
        } );
    for( auto& t : vT )
        t.join( );
    double fSec = std::chrono::duration<double>( std::chrono::steady_clock::now( ) - tStart ).count( );
    printf( "{\"calls\": %zu, \"threads\": %d, \"seconds\": %.6f}\n", uiTotal, iThreads, fSec );
    return 0;
}

int main( int argc, char** argv )
{
    if( argc < 2 )
    {
        std::cerr << "usage: ref_dump index|align|ksw|bench|kswbench ..." << std::endl;
        return 2;
    }
    std::string sCmd = argv[ 1 ];
    try
    {
        if( sCmd == "index" && argc >= 4 )
            return cmdIndex( argc, argv );
        if( sCmd == "align" && argc >= 6 )
            return cmdAlign( argc, argv );
        if( sCmd == "reads" && argc >= 3 )
            return cmdReads( argc, argv );
        if( sCmd == "sam" && argc >= 6 )
            return cmdSam( argc, argv );
        if( sCmd == "ksw" && argc >= 4 )
            return cmdKsw( argc, argv );
        if( sCmd == "bench" && argc >= 6 )
            return cmdBench( argc, argv );
        if( sCmd == "kswbench" && argc >= 5 )
            return cmdKswBench( argc, argv );
    }
    catch( std::exception& e )
    {
        std::cerr << "ref_dump: exception: " << e.what( ) << std::endl;
        return 1;
    }
    std::cerr << "ref_dump: bad arguments" << std::endl;
    return 2;
}
