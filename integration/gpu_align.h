/* Adapter modules that bind libma_b200.so from INSIDE the reference tree (ITBE-Lab/ma): compiled against the
 * reference's own headers (libs/ms/inc, libs/ma/inc) and linked with its library, with the reference's module
 * signatures (libs/ms/inc/ms/module/module.h:63-122), so that they can be wired into its graphs like the CPU modules
 * (libs/ma/src/util/export.cpp:72-202). This is the file a maintainer would add as libs/ma/inc/ma/module/gpuAlign.h.
 * Built and run against the unmodified reference by integration/Makefile + tests/test_reference_boundary_gpu.py.
 *
 *   GpuIndex          FMIndex + Pack of a genome, uploaded once into the HBM of one device (ma_b200_index_upload)
 *   GpuAlign          batch module: ContainerVector<NucSeq> -> per read the MappingQuality result vector
 *                     (BinarySeeding .. NeedlemanWunsch, MappingQuality and, for paired presets, PairedReads on the GPU)
 *   GpuAlignPerRead   per-read module with the signature of the whole CPU chain: execute( NucSeq ) blocks until the
 *                     batch it joined has been aligned (SURVEY.md §8(b) option (i)), so that per-read graphs
 *                     (setUpCompGraph: one graph per thread, export.cpp:85-127) run unmodified
 */
#pragma once
#include "ma/container/alignment.h"
#include "ma/container/fMIndex.h"
#include "ma/container/nucSeq.h"
#include "ma/container/pack.h"
#include "ma/module/fileReader.h"
#include "ma/module/fileWriter.h"
#include "ms/module/splitter.h"
#include "ma_b200.h"
#include "ms/container/container.h"
#include "ms/module/module.h"
#include "ms/util/parameter.h"
#include <chrono>
#include <condition_variable>
#include <fstream>
#include <mutex>

namespace libMA
{

/* ma_b200_params from the selected presetting and pGlobalParams, field by field where the CPU modules copy their const
 * members (binarySeeding.h:559-569, stripOfConsideration.h:179-190, harmonization.h:156-173, needlemanWunsch.h:91-103,
 * mappingQuality.h:26-36, pairedReads.h:43-55) */
inline ma_b200_params gpuParams( const ParameterSetManager& rParameters )
{
    ma_b200_params p;
    ma_b200_params_preset( "default", &p );
    auto s = rParameters.getSelected( );
    p.match = pGlobalParams->iMatch->get( ), p.mismatch = pGlobalParams->iMissMatch->get( );
    p.gap = pGlobalParams->iGap->get( ), p.extend = pGlobalParams->iExtend->get( );
    p.gap2 = pGlobalParams->iGap2->get( ), p.extend2 = pGlobalParams->iExtend2->get( );
    p.sv_penalty = pGlobalParams->uiSVPenalty->get( );
    p.seeding_technique = s->xSeedingTechnique->get( ) == "maxSpan" ? 0 : 1;
    p.min_seed_length = s->xMinSeedLength->get( );
    p.min_ambiguity = s->xMinimalSeedAmbiguity->get( ), p.max_ambiguity = s->xMaximalSeedAmbiguity->get( );
    p.seed_drop_min_size = s->xMinimalSeedSizeDrop->get( ), p.seed_drop_factor = s->xRelMinSeedSizeAmount->get( );
    p.max_num_soc = s->xMaxNumSoC->get( ), p.min_num_soc = s->xMinNumSoC->get( );
    p.soc_width = s->xSoCWidth->get( ), p.rectangular_soc = s->xRectangularSoc->get( ) ? 1 : 0;
    p.soc_score_drop = s->xSoCScoreDecreaseTolerance->get( );
    p.harm_score_min = s->xHarmScoreMin->get( ), p.harm_score_min_rel = s->xHarmScoreMinRel->get( );
    p.score_diff_tolerance = s->xScoreDiffTolerance->get( );
    p.max_score_lookahead = s->xMaxScoreLookahead->get( ), p.switch_qlen = s->xSwitchQlen->get( );
    p.max_delta_dist = s->xMaxDeltaDist->get( ), p.min_delta_dist = s->xMinDeltaDist->get( );
    p.optimistic_gap_estimation = s->xOptimisticGapCostEstimation->get( ) ? 1 : 0;
    p.gap_cost_cutting = s->xDisableGapCostEstimationCutting->get( ) ? 0 : 1;
    p.max_gap_area = s->xMaxGapArea->get( );
    p.genome_size_disable = s->xGenomeSizeDisable->get( ), p.disable_heuristics = s->xDisableHeuristics->get( ) ? 1 : 0;
    p.padding = s->xPadding->get( ), p.bandwidth_ext = s->xBandwidthDPExtension->get( );
    p.min_bandwidth_gap = s->xMinBandwidthGapFilling->get( ), p.zdrop = s->xZDrop->get( );
    p.report_n = s->xReportN->get( ), p.min_alignment_score = s->xMinAlignmentScore->get( );
    p.max_supplementary_per_prim = s->xMaxSupplementaryPerPrim->get( );
    p.max_overlap_supplementary = s->xMaxOverlapSupplementary->get( );
    p.use_paired_reads = s->xUsePairedReads->get( ) ? 1 : 0;
    p.paired_mean = s->xMeanPairedReadDistance->get( ), p.paired_std = s->xStdPairedReadDistance->get( );
    p.paired_bonus = s->xPairedBonus->get( );
    return p;
}

/* The index of one genome resident on one GPU. A subclass of FMIndex: bwt / sa are protected members of the reference's
 * class (fMIndex.h:213-230) and go up exactly as they are; the packed forward strand is read from <prefix>.pac (its
 * vector is private in Pack, pack.h:151), the contig table comes from Pack's public accessors. */
class GpuIndex : public FMIndex
{
  public:
    ma_b200_ctx* pCtx = nullptr;
    std::shared_ptr<Pack> pPack;

    GpuIndex( int iDevice, const std::string& sPrefix ) : FMIndex( ), pPack( std::make_shared<Pack>( ) )
    {
        this->vLoadFMIndex( sPrefix );
        pPack->vLoadCollection( sPrefix );
        if( ma_b200_create( iDevice, &pCtx ) != MA_B200_OK )
            throw std::runtime_error( "GpuIndex: no CUDA device (the GPU path has no CPU fallback)" );
        std::ifstream xPac( sPrefix + ".pac", std::ios::binary | std::ios::ate );
        if( !xPac )
            throw std::runtime_error( "GpuIndex: cannot open " + sPrefix + ".pac" );
        std::vector<uint8_t> vPac( (size_t)xPac.tellg( ) );
        xPac.seekg( 0 );
        xPac.read( (char*)vPac.data( ), (std::streamsize)vPac.size( ) );
        const int64_t iFwd = (int64_t)pPack->uiUnpackedSizeForwardStrand;
        std::vector<int64_t> vStart, vLen;
        for( int64_t i = 0; i < (int64_t)pPack->uiNumContigs( ); i++ )
        {
            vStart.push_back( (int64_t)pPack->startOfSequenceWithId( i ) );
            vLen.push_back( (int64_t)pPack->lengthOfSequenceWithId( i ) );
        }
        const int64_t aL2[ 5 ] = { 0, (int64_t)L2[ 1 ], (int64_t)L2[ 2 ], (int64_t)L2[ 3 ], (int64_t)L2[ 4 ] };
        check( ma_b200_index_upload( pCtx, bwt.data( ), (int64_t)bwt.size( ), aL2, (int64_t)primary, (int64_t)uiRefSeqLength,
                                     (const int64_t*)sa.data( ), (int64_t)sa.size( ), sa_intv, vPac.data( ),
                                     ( iFwd + 3 ) / 4, iFwd, vStart.data( ), vLen.data( ), (int32_t)vStart.size( ) ) );
    }
    GpuIndex( const GpuIndex& ) = delete;
    ~GpuIndex( )
    {
        ma_b200_destroy( pCtx );
    }
    /* a non-zero code of the C ABI becomes the std::runtime_error the graph runtime collects (module.h:339-377) */
    void check( int iRc ) const
    {
        if( iRc != MA_B200_OK )
            throw std::runtime_error( std::string( "libma_b200: " ) + ma_b200_last_error( pCtx ) );
    }
};

typedef libMS::ContainerVector<std::shared_ptr<Alignment>> AlignmentVector;

/* One call of ma_b200_align_batch for a vector of reads; element i of the result is what
 * MappingQuality::execute( query i, NeedlemanWunsch::execute( ... ) ) returns on the CPU (mappingQuality.cpp:11-131);
 * with "Use Paired Reads" the reads 2k, 2k + 1 are mates and element k of vPaired is PairedReads::execute's vector
 * (pairedReads.cpp:15-121) holding the same Alignment objects. */
class GpuAlign : public libMS::Module<libMS::ContainerVector<std::shared_ptr<AlignmentVector>>, false,
                                      libMS::ContainerVector<std::shared_ptr<NucSeq>>, GpuIndex>
{
  public:
    const ma_b200_params xParams;
    uint32_t uiSrandBase = 0; /* RANSAC stream of read i: srand( uiSrandBase + i ), SURVEY.md A-5 */
    bool bSrand = false;
    std::vector<std::shared_ptr<AlignmentVector>> vPaired; /* filled by execute for paired presets */

    GpuAlign( const ParameterSetManager& rParameters ) : xParams( gpuParams( rParameters ) )
    {}

    virtual std::shared_ptr<libMS::ContainerVector<std::shared_ptr<AlignmentVector>>>
    execute( std::shared_ptr<libMS::ContainerVector<std::shared_ptr<NucSeq>>> pReads, std::shared_ptr<GpuIndex> pIdx )
    {
        const int64_t n = (int64_t)pReads->size( );
        std::vector<int64_t> vOff( (size_t)n + 1, 0 );
        for( int64_t i = 0; i < n; i++ )
            vOff[ i + 1 ] = vOff[ i ] + (int64_t)( *pReads )[ i ]->length( );
        std::vector<uint8_t> vSlab( (size_t)vOff[ n ] + 1 );
        for( int64_t i = 0; i < n; i++ )
            memcpy( vSlab.data( ) + vOff[ i ], ( *pReads )[ i ]->pGetSequenceRef( ), ( *pReads )[ i ]->length( ) );
        ma_b200_params p = xParams;
        if( bSrand )
            p.srand_base = uiSrandBase;
        pIdx->check( ma_b200_set_params( pIdx->pCtx, &p ) );
        /* upload / run / download: the record arrays are sized from the counts of the run */
        ma_b200_align_stats xStats;
        pIdx->check( ma_b200_align_upload( pIdx->pCtx, n, vSlab.data( ), vOff.data( ) ) );
        pIdx->check( ma_b200_align_run( pIdx->pCtx, MA_B200_STAGE_MAPQ, 0, &xStats ) );
        std::vector<ma_b200_read_info> vInfo( (size_t)n + 1 );
        std::vector<ma_b200_alignment> vAlns( (size_t)xStats.n_sets + 1 );
        std::vector<uint32_t> vRuns( (size_t)xStats.n_runs + 1 );
        pIdx->check( ma_b200_align_download( pIdx->pCtx, vInfo.data( ), vAlns.data( ), (int64_t)vAlns.size( ), vRuns.data( ),
                                             (int64_t)vRuns.size( ) ) );
        auto pRet = std::make_shared<libMS::ContainerVector<std::shared_ptr<AlignmentVector>>>( );
        vPaired.clear( );
        for( int64_t i = 0; i < n; i++ )
        {
            if( vInfo[ i ].status )
                std::cerr << "GpuAlign: read " << ( *pReads )[ i ]->sName << " exceeds a capacity of the GPU path (status "
                          << vInfo[ i ].status << "), reported as unaligned" << std::endl;
            /* records of the read in MappingQuality's order (rank_mq), those it dropped (-1) left out */
            std::vector<const ma_b200_alignment*> vOrd;
            for( int k = 0; k < vInfo[ i ].n_sets; k++ )
                if( vAlns[ vInfo[ i ].set_off + k ].rank_mq >= 0 )
                    vOrd.push_back( &vAlns[ vInfo[ i ].set_off + k ] );
            std::sort( vOrd.begin( ), vOrd.end( ),
                       []( const ma_b200_alignment* a, const ma_b200_alignment* b ) { return a->rank_mq < b->rank_mq; } );
            auto pVec = std::make_shared<AlignmentVector>( );
            std::vector<std::pair<int, std::shared_ptr<Alignment>>> vPairOrd;
            for( const ma_b200_alignment* r : vOrd )
            {
                auto pA = std::make_shared<Alignment>( (nucSeqIndex)r->begin_ref );
                pA->uiEndOnRef = (nucSeqIndex)r->end_ref, pA->uiBeginOnQuery = (nucSeqIndex)r->begin_q;
                pA->uiEndOnQuery = (nucSeqIndex)r->end_q, pA->uiLength = (nucSeqIndex)r->length;
                pA->iScore = r->score, pA->fMappingQuality = r->mapq;
                pA->bSecondary = ( r->flags & MA_B200_ALN_SECONDARY ) != 0;
                pA->bSupplementary = ( r->flags & MA_B200_ALN_SUPPLEMENTARY ) != 0;
                pA->xStats.index_of_strip = r->soc_index;
                pA->xStats.sName = ( *pReads )[ i ]->sName;
                pA->xStats.bFirst = xParams.use_paired_reads ? !( i & 1 ) : ( r->flags & MA_B200_ALN_FIRST_MATE ) != 0;
                /* MatchType numbering is the reference's (alignment.h:40-47): seed 0, match 1, missmatch 2, ins 3, del 4 */
                for( int k = 0; k < r->n_runs; k++ )
                {
                    const uint32_t w = vRuns[ (size_t)r->run_off + k ];
                    pA->data.emplace_back( (MatchType)( w & 7 ), (nucSeqIndex)( w >> 3 ) );
                }
                pVec->push_back( pA );
                if( r->pair_rank >= 0 )
                    vPairOrd.emplace_back( r->pair_rank, pA );
            }
            pRet->push_back( pVec );
            if( xParams.use_paired_reads && !( i & 1 ) )
                vPendingMate = vPairOrd;
            else if( xParams.use_paired_reads )
            { /* PairedReads' vector of the pair (i - 1, i): the chosen pair, its two alignments linked
                 (pairedReads.cpp:94-95, 117-121), or one mate's vector passed through (:28-31) */
                std::vector<std::pair<int, std::shared_ptr<Alignment>>> vBoth = vPendingMate;
                vBoth.insert( vBoth.end( ), vPairOrd.begin( ), vPairOrd.end( ) );
                std::sort( vBoth.begin( ), vBoth.end( ), []( const auto& a, const auto& b ) { return a.first < b.first; } );
                auto pPair = std::make_shared<AlignmentVector>( );
                for( auto& rP : vBoth )
                    pPair->push_back( rP.second );
                if( !( *pRet )[ (size_t)i - 1 ]->empty( ) && !pVec->empty( ) && pPair->size( ) == 2 )
                {
                    ( *pPair )[ 0 ]->xStats.pOther = std::weak_ptr<Alignment>( ( *pPair )[ 1 ] );
                    ( *pPair )[ 1 ]->xStats.pOther = std::weak_ptr<Alignment>( ( *pPair )[ 0 ] );
                }
                vPaired.push_back( pPair );
                vPendingMate.clear( );
            }
        }
        return pRet;
    }

  private:
    std::vector<std::pair<int, std::shared_ptr<Alignment>>> vPendingMate;
};

/* Per-read front of GpuAlign for the reference's one-graph-per-thread model (SURVEY.md §8(b) option (i)): every graph
 * thread calls execute( read, index ) and blocks; the call that completes a batch of uiBatch reads runs the batch on the
 * GPU and wakes the others. A graph thread cannot tell the module that it has run out of reads, so a waiting call that
 * sees no progress for uiPatienceMicroseconds flushes what is open (the last, short batches of a run). One object is
 * shared by all graphs, like the reference's shared FileWriter, which serialises itself with a mutex
 * (fileWriter.h:386-398). */
class GpuAlignPerRead : public libMS::Module<AlignmentVector, false, NucSeq, GpuIndex>
{
    struct Slot
    {
        std::shared_ptr<NucSeq> pQuery;
        std::shared_ptr<AlignmentVector> pResult;
        bool bDone = false;
        std::string sError;
    };
    GpuAlign xBatch;
    const size_t uiBatch;
    const unsigned int uiPatienceMicroseconds;
    std::mutex xMutex;
    std::condition_variable xCv;
    std::vector<std::shared_ptr<Slot>> vOpen;
    uint64_t uiReadCounter = 0;

    void flush( std::shared_ptr<GpuIndex> pIdx )
    { /* with the lock held: one context, one GPU call at a time; reads that arrive meanwhile wait for the lock */
        std::vector<std::shared_ptr<Slot>> vMine;
        vMine.swap( vOpen );
        const uint64_t uiFirst = uiReadCounter;
        uiReadCounter += vMine.size( );
        auto pReads = std::make_shared<libMS::ContainerVector<std::shared_ptr<NucSeq>>>( );
        for( auto& pS : vMine )
            pReads->push_back( pS->pQuery );
        std::string sError;
        std::shared_ptr<libMS::ContainerVector<std::shared_ptr<AlignmentVector>>> pRes;
        try
        {
            if( xBatch.bSrand )
                xBatch.uiSrandBase = (uint32_t)( uiSrandFirst + uiFirst );
            pRes = xBatch.execute( pReads, pIdx );
        }
        catch( const std::exception& e )
        {
            sError = e.what( );
        }
        for( size_t i = 0; i < vMine.size( ); i++ )
        {
            vMine[ i ]->sError = sError;
            if( pRes )
                vMine[ i ]->pResult = ( *pRes )[ i ];
            vMine[ i ]->bDone = true;
        }
        xCv.notify_all( );
    }

  public:
    uint64_t uiSrandFirst = 0;
    GpuAlignPerRead( const ParameterSetManager& rParameters, size_t uiBatch, unsigned int uiPatienceMicroseconds = 2000 )
        : xBatch( rParameters ), uiBatch( uiBatch ), uiPatienceMicroseconds( uiPatienceMicroseconds )
    {}
    void setSrand( uint64_t uiBase )
    {
        xBatch.bSrand = true, uiSrandFirst = uiBase;
    }
    virtual std::shared_ptr<AlignmentVector> execute( std::shared_ptr<NucSeq> pQuery, std::shared_ptr<GpuIndex> pIdx )
    {
        auto pSlot = std::make_shared<Slot>( );
        pSlot->pQuery = pQuery;
        std::unique_lock<std::mutex> xLock( xMutex );
        vOpen.push_back( pSlot );
        if( vOpen.size( ) >= uiBatch )
            flush( pIdx );
        while( !pSlot->bDone )
            if( !xCv.wait_for( xLock, std::chrono::microseconds( uiPatienceMicroseconds ), [ & ] { return pSlot->bDone; } ) )
                if( !pSlot->bDone ) /* nobody completed the batch in time: the other graphs are busy or out of reads */
                    flush( pIdx );
        if( !pSlot->sError.empty( ) )
            throw std::runtime_error( pSlot->sError );
        return pSlot->pResult;
    }
};

/* setUpCompGraph (libs/ma/src/util/export.cpp:72-128) with the five CPU modules BinarySeeding .. MappingQuality replaced
 * by ONE shared GpuAlignPerRead: the same file-stream queue, lock, FileReader, writer, progress printer and unlock
 * modules, one graph per thread, evaluated by BasePledge::simultaneousGet exactly like ExecutionContext::doAlign does
 * (execution-context.h:291-406) — a GPU build of doAlign calls this instead of setUpCompGraph. */
/* the writer interface of setUpCompGraph (TP_WRITER, ma/util/export.h:166-167; that header also pulls in the minimizer
 * modules, which are not part of this build) */
typedef libMS::Module<libMS::Container, false, NucSeq, libMS::ContainerVector<std::shared_ptr<Alignment>>, Pack> TP_GPU_WRITER;

inline std::vector<std::shared_ptr<libMS::BasePledge>>
setUpCompGraphGpu( const ParameterSetManager& rParameters, std::shared_ptr<libMS::Pledge<Pack>> pPack,
                   std::shared_ptr<libMS::Pledge<GpuIndex>> pGpuIndex,
                   std::shared_ptr<libMS::Pledge<FileStreamQueue, false>> pQueue, std::shared_ptr<TP_GPU_WRITER> pWriter,
                   std::shared_ptr<GpuAlignPerRead> pGpuAlign, unsigned int uiThreads )
{
    using namespace libMS;
    auto pFileStreamPicker = std::make_shared<QueuePicker<FileStream>>( rParameters );
    auto pLock = std::make_shared<Lock<FileStream>>( rParameters );
    auto pFileReader = std::make_shared<FileReader>( rParameters );
    auto pFileStreamPlacer = std::make_shared<QueuePlacer<NucSeq, FileStream>>( rParameters );
    auto pProgressPrinter = std::make_shared<ProgressPrinter<FileStreamQueue>>( rParameters );
    std::vector<std::shared_ptr<BasePledge>> aRet;
    BasePledge::parallelGraph( uiThreads, [ & ]( ) {
        auto pPickedFile = promiseMe( pFileStreamPicker, pQueue );
        auto pLockedFile = promiseMe( pLock, pPickedFile );
        auto pQuery_ = promiseMe( pFileReader, pLockedFile );
        auto pQuery = promiseMe( pFileStreamPlacer, pQuery_, pLockedFile, pQueue );
        auto pAlignmentsWQuality = promiseMe( pGpuAlign, pQuery, pGpuIndex ); /* seeding .. MappingQuality on the GPU */
        auto pEmptyContainer = promiseMe( pWriter, pQuery, pAlignmentsWQuality, pPack );
        auto pEmptyContainer_ = promiseMe( pProgressPrinter, pEmptyContainer, pQueue );
        auto pUnlockResult =
            promiseMe( std::make_shared<UnLock<libMS::Container>>( rParameters, pLockedFile ), pEmptyContainer_ );
        aRet.push_back( pUnlockResult );
    } );
    return aRet;
}

} // namespace libMA
