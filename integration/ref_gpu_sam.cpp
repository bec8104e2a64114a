// Runs the adapter modules of integration/gpu_align.h inside the UNMODIFIED reference: its FileReader / NucSeq, its
// ParameterSetManager presets, its Pack and its FileWriter / PairedFileWriter, with GpuAlign in place of the module
// chain of setUpCompGraph (libs/ma/src/util/export.cpp:99-126, 160-199). Linked with oracle/_ref/libma_ref.so (the
// compiled reference) and libma_b200.so. Test infrastructure for the boundary, not part of the product.
//   ref_gpu_sam batch   <index prefix> <reads> <preset> <out.sam> <srand base>
//   ref_gpu_sam perread <index prefix> <reads> <preset> <out.sam> <srand base> <threads> <batch>
//   ref_gpu_sam graph   <index prefix> <reads.fa|fq> <preset> <out.sam> <srand base> <threads> <batch>
//       the reference's computational graph itself: setUpCompGraphGpu + BasePledge::simultaneousGet, i.e. what
//       ExecutionContext::doAlign does with the GPU module in place of the five CPU modules
#include "gpu_align.h"
#include "ma/module/fileReader.h"
#include "ma/module/fileWriter.h"
#include <thread>

using namespace libMA;
using namespace libMS;

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
    std::string sLine; // one read per line, named r<i> (the format of tests/golden/gold_reads_*.txt)
    while( std::getline( xProbe, sLine ) )
        if( !sLine.empty( ) )
        {
            vRet.push_back( std::make_shared<NucSeq>( sLine ) );
            vRet.back( )->sName = "r" + std::to_string( vRet.size( ) - 1 );
        }
    return vRet;
}

int main( int argc, char** argv )
{
    if( argc < 7 )
    {
        std::cerr << "usage: ref_gpu_sam batch|perread <prefix> <reads> <preset> <out.sam> <srand> [threads batch]\n";
        return 2;
    }
    try
    {
        const std::string sMode = argv[ 1 ], sPrefix = argv[ 2 ];
        ParameterSetManager xP;
        std::string sPreset = argv[ 4 ];
        for( auto& c : sPreset )
            c = std::tolower( c );
        xP.setSelected( sPreset );
        auto vReads = sMode == "graph" ? std::vector<std::shared_ptr<NucSeq>>( ) : readQueries( xP, argv[ 3 ] );
        auto pIdx = std::make_shared<GpuIndex>( 0, sPrefix );
        auto pPack = pIdx->pPack;
        const bool bPaired = xP.getSelected( )->xUsePairedReads->get( );
        const uint32_t uiSrand = (uint32_t)atoll( argv[ 6 ] );
        if( sMode == "batch" )
        { // FileReader -> batch -> GpuAlign -> FileWriter
            auto pBatch = std::make_shared<ContainerVector<std::shared_ptr<NucSeq>>>( );
            for( auto& pQ : vReads )
                pBatch->push_back( pQ );
            GpuAlign xAlign( xP );
            xAlign.bSrand = true, xAlign.uiSrandBase = uiSrand;
            auto pRes = xAlign.execute( pBatch, pIdx );
            if( bPaired )
            {
                PairedFileWriter xW( xP, std::string( argv[ 5 ] ), pPack );
                for( size_t i = 0; i + 1 < vReads.size( ); i += 2 )
                    xW.execute( vReads[ i ], vReads[ i + 1 ], xAlign.vPaired[ i / 2 ], pPack );
            }
            else
            {
                FileWriter xW( xP, std::string( argv[ 5 ] ), pPack );
                for( size_t i = 0; i < vReads.size( ); i++ )
                    xW.execute( vReads[ i ], ( *pRes )[ i ], pPack );
            }
            return 0;
        }
        if( sMode == "graph" )
        {
            const unsigned int uiThreads = argc > 7 ? (unsigned int)atoi( argv[ 7 ] ) : 4;
            const size_t uiBatch = argc > 8 ? (size_t)atoi( argv[ 8 ] ) : 64;
            auto pAlign = std::make_shared<GpuAlignPerRead>( xP, uiBatch );
            pAlign->setSrand( uiSrand );
            auto pPackPledge = std::make_shared<Pledge<Pack>>( );
            pPackPledge->set( pPack );
            auto pIdxPledge = std::make_shared<Pledge<GpuIndex>>( );
            pIdxPledge->set( pIdx );
            auto pInitVec = std::make_shared<ContainerVector<std::shared_ptr<FileStream>>>( );
            pInitVec->push_back( std::make_shared<FileStreamFromPath>( std::string( argv[ 3 ] ) ) );
            auto pQueuePledge = std::make_shared<Pledge<FileStreamQueue>>( );
            pQueuePledge->set( std::make_shared<FileStreamQueue>( pInitVec ) );
            std::shared_ptr<TP_GPU_WRITER> pWriter( new FileWriter( xP, std::string( argv[ 5 ] ), pPack ) );
            auto aGraphSinks = setUpCompGraphGpu( xP, pPackPledge, pIdxPledge, pQueuePledge, pWriter, pAlign, uiThreads );
            BasePledge::simultaneousGet( aGraphSinks );
            return 0;
        }
        // per-read graphs: T threads, each the loop of one computational graph of setUpCompGraph (reader -> aligner ->
        // writer); the aligner module is shared and batches behind the scenes. Output order is the threads' order, as
        // with the reference's own shared FileWriter.
        const size_t uiThreads = argc > 7 ? (size_t)atoi( argv[ 7 ] ) : 4, uiBatch = argc > 8 ? (size_t)atoi( argv[ 8 ] ) : 64;
        auto pAlign = std::make_shared<GpuAlignPerRead>( xP, uiBatch );
        auto pW = std::make_shared<FileWriter>( xP, std::string( argv[ 5 ] ), pPack );
        std::mutex xNext;
        size_t uiNext = 0;
        std::vector<std::string> vErr( uiThreads );
        std::vector<std::thread> vT;
        // reads are handed out in order and their RANSAC stream follows the order in which they join batches, which with
        // several threads is not the file order: exactness against the golden SAM is checked with one thread, the
        // multi-threaded run for completeness (every read written once).
        pAlign->setSrand( uiSrand );
        for( size_t t = 0; t < uiThreads; t++ )
            vT.emplace_back( [ &, t ]( ) {
                try
                {
                    while( true )
                    {
                        std::shared_ptr<NucSeq> pQ;
                        {
                            std::lock_guard<std::mutex> g( xNext );
                            if( uiNext < vReads.size( ) )
                                pQ = vReads[ uiNext++ ];
                        }
                        if( !pQ )
                            break;
                        pW->execute( pQ, pAlign->execute( pQ, pIdx ), pPack );
                    }
                }
                catch( const std::exception& e )
                {
                    vErr[ t ] = e.what( );
                }
            } );
        for( auto& t : vT )
            t.join( );
        for( auto& e : vErr )
            if( !e.empty( ) )
                throw std::runtime_error( e );
        return 0;
    }
    catch( const std::exception& e )
    {
        std::cerr << "ref_gpu_sam: " << e.what( ) << std::endl;
        return 1;
    }
}
