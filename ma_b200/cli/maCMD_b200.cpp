// maCMD_b200 — command-line front end of the B200 alignment path with the reference CLI's options for that path
// (cmdMa.cpp:252-431): reads FASTA/FASTQ like FileReader, aligns batches of reads on the GPU through
// the C ABI (libma_b200.so) and writes the SAM text of the reference's FileWriter / PairedFileWriter.
//
//   maCMD_b200 -x <index prefix> -i <reads.fq[,more.fq]> [-m <mates.fq[,more.fq]>] [-o <out.sam>] [-p <presetting>]
//
//   -x, --Index        prefix of the reference's index files (.bwt .sa .pac .ann .amb), as written by maCMD --Create_Index
//   -X, --Create_Index <fasta>,<folder>,<name>   builds the index on the GPU and writes the reference's index files
//                      (genomes below 2^30 bases without N); -x also takes the <name>.json written here
//   -i, --In           FASTA / FASTQ file(s), plain or gzip-compressed, comma separated
//   -m, --MateIn       mate file(s); switches "Use Paired Reads" on like the reference (cmdMa.cpp:323-330)
//   -o, --Out          SAM file (default: standard output)
//   -p, --Presetting   Default | Illumina | Illumina_Paired | PacBio | Nanopore (default: Default)
//   -t                 host threads that format SAM records (default: all but three; the alignment itself runs on the GPU)
//   --Verbose          prints the busy time of the host stages
//   --Detect_Small_Inversions [true|false], --Z_Drop_Inversions <n>   the reference's parameters of that name:
//                      SmallInversions between MappingQuality and the writer / PairedReads (then paired on the host)
//   --Use_M_in_CIGAR [true|false], --Soft_clip, --Omit_Secondary_Alignments, --Omit_Supplementary_Alignments
//                      the writers' flags of the reference (default: M CIGARs, hard clips, everything written)
//   --Interleaved      with a paired presetting and no -m: reads 2k, 2k+1 of -i are mates (not in the reference)
//   --Devices <a,b,..> CUDA devices (default 0). The index is replicated on every device, batches go to whichever device
//                      is free, no collective is involved; the output order does not depend on it
//   --Batch <n>        reads per GPU batch (default 500000; pairs are never split)
//   --Srand <n>        RANSAC stream of read i is srand(n + i) (parity contract; default 0)
//
// Records are written in input order. There is no CPU path: the program fails when no CUDA device is present.
#include "../../include/ma_b200_modules.hpp"
#include "../../include/ma_b200_sam.hpp"
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <exception>
#include <fstream>
#include <functional>
#include <iostream>
#include <map>
#include <mutex>
#include <thread>

using namespace libMA_b200;

static std::vector<std::string> splitList( const std::string& s )
{
    std::vector<std::string> v;
    size_t b = 0;
    while( b <= s.size( ) )
    {
        size_t e = s.find( ',', b );
        if( e == std::string::npos )
            e = s.size( );
        if( e > b )
            v.push_back( s.substr( b, e - b ) );
        b = e + 1;
    }
    return v;
}

static std::string lower( std::string s )
{
    for( auto& c : s )
        c = (char)tolower( c );
    return s;
}

// hand-over between the host stages; push() returns false once the queue is closed
template <typename T> class BoundedQueue
{
    std::mutex xMutex;
    std::condition_variable xCv;
    std::deque<T> vItems;
    size_t uiCapacity;
    bool bClosed = false;

  public:
    explicit BoundedQueue( size_t uiCap ) : uiCapacity( uiCap )
    {}
    bool push( T x )
    {
        std::unique_lock<std::mutex> xLock( xMutex );
        xCv.wait( xLock, [ & ]( ) { return bClosed || vItems.size( ) < uiCapacity; } );
        if( bClosed )
            return false;
        vItems.push_back( std::move( x ) );
        xCv.notify_all( );
        return true;
    }
    bool pop( T& x ) // false: closed and drained
    {
        std::unique_lock<std::mutex> xLock( xMutex );
        xCv.wait( xLock, [ & ]( ) { return bClosed || !vItems.empty( ); } );
        if( vItems.empty( ) )
            return false;
        x = std::move( vItems.front( ) );
        vItems.pop_front( );
        xCv.notify_all( );
        return true;
    }
    void close( )
    {
        std::lock_guard<std::mutex> xLock( xMutex );
        bClosed = true;
        xCv.notify_all( );
    }
};

// one stream of reads over several files
class ReadStream
{
    std::vector<std::string> vFiles;
    size_t uiFile = 0;
    std::shared_ptr<ReadParser> pParser;

  public:
    // byte range of one record in the file of pOwner (converted later, possibly on another thread)
    struct Record
    {
        const ReadParser* pOwner;
        size_t uiBegin, uiEnd;
    };
    explicit ReadStream( std::vector<std::string> v ) : vFiles( std::move( v ) )
    {}
    // serial pass: the next record; vKeepAlive collects the parsers whose records are still to be converted
    bool next( Record& r, std::vector<std::shared_ptr<ReadParser>>& vKeepAlive )
    {
        while( true )
        {
            if( !pParser )
            {
                if( uiFile >= vFiles.size( ) )
                    return false;
                pParser = std::make_shared<ReadParser>( vFiles[ uiFile++ ] );
            }
            if( pParser->nextRecord( r.uiBegin, r.uiEnd ) )
            {
                r.pOwner = pParser.get( );
                if( vKeepAlive.empty( ) || vKeepAlive.back( ) != pParser )
                    vKeepAlive.push_back( pParser );
                return true;
            }
            pParser.reset( );
        }
    }
};

// --Create_Index <fasta>,<folder>,<name> (cmdMa.cpp:332-345, ExecutionContext::makeIndexAndPackForGenome,
// execution-context.h:108-138): <folder>/<fasta stem>.{bwt,sa,pac,ann,amb} in the reference's file formats (fMIndex.h:
// 515-599, pack.h:275-451) and <folder>/<name>.json. The index is built on the GPU (ma_b200_index_build, bit-identical
// to the reference's FMIndex( pPack )) for genomes below 2^30 bases without N.
static void createIndex( const std::string& sFasta, const std::string& sFolder, const std::string& sTitle, int iDevice )
{
    std::ifstream xIn( sFasta );
    if( !xIn )
        throw std::runtime_error( "Unable to open file " + sFasta );
    std::vector<uint8_t> vFwd;
    std::vector<std::string> vNames, vComments;
    std::vector<int64_t> vStart, vLength;
    std::string sLine;
    while( std::getline( xIn, sLine ) )
    {
        while( !sLine.empty( ) && ( sLine.back( ) == '\r' || sLine.back( ) == ' ' ) )
            sLine.pop_back( );
        if( sLine.empty( ) )
            continue;
        if( sLine[ 0 ] == '>' )
        {
            const size_t uiBlank = sLine.find( ' ' );
            vNames.push_back( sLine.substr( 1, uiBlank == std::string::npos ? std::string::npos : uiBlank - 1 ) );
            vComments.push_back( uiBlank == std::string::npos ? "" : sLine.substr( uiBlank + 1 ) );
            vStart.push_back( (int64_t)vFwd.size( ) ), vLength.push_back( 0 );
            continue;
        }
        if( vNames.empty( ) )
            throw std::runtime_error( sFasta + " is not a FASTA file" );
        for( char c : sLine )
        {
            const int b = c == 'A' || c == 'a' ? 0 : c == 'C' || c == 'c' ? 1 : c == 'G' || c == 'g' ? 2
                                                   : c == 'T' || c == 't' ? 3 : -1;
            if( b < 0 )
                throw std::runtime_error( "--Create_Index: '" + std::string( 1, c ) + "' in " + vNames.back( ) +
                                          ": only genomes without N / IUPAC codes can be indexed here (the reference "
                                          "replaces them by random bases and records holes)" );
            vFwd.push_back( (uint8_t)b );
        }
        vLength.back( ) = (int64_t)vFwd.size( ) - vStart.back( );
    }
    if( vFwd.empty( ) )
        throw std::runtime_error( "--Create_Index: no sequence in " + sFasta );
    ma_b200_ctx* pCtx = nullptr;
    if( ma_b200_create( iDevice, &pCtx ) != MA_B200_OK )
        throw std::runtime_error( "ma_b200_create failed: no CUDA device (there is no CPU fallback)" );
    auto check = [ & ]( int rc ) {
        if( rc != MA_B200_OK )
        {
            const std::string sErr = ma_b200_last_error( pCtx );
            ma_b200_destroy( pCtx );
            throw std::runtime_error( "ma_b200: " + sErr );
        }
    };
    check( ma_b200_index_build( pCtx, vFwd.data( ), (int64_t)vFwd.size( ), vStart.data( ), vLength.data( ),
                                (int32_t)vNames.size( ) ) );
    int64_t nWords = 0, nSa = 0, nPac = 0, iPrimary = 0, aL2[ 5 ] = { 0, 0, 0, 0, 0 };
    check( ma_b200_index_sizes( pCtx, &nWords, &nSa, &nPac, &iPrimary, aL2 ) );
    std::vector<uint32_t> vBwt( (size_t)nWords );
    std::vector<int64_t> vSa( (size_t)nSa );
    std::vector<uint8_t> vPac( (size_t)nPac );
    check( ma_b200_index_download( pCtx, vBwt.data( ), vSa.data( ), vPac.data( ) ) );
    ma_b200_destroy( pCtx );

    std::string sStem = sFasta.substr( sFasta.find_last_of( '/' ) == std::string::npos ? 0 : sFasta.find_last_of( '/' ) + 1 );
    if( sStem.find_last_of( '.' ) != std::string::npos )
        sStem = sStem.substr( 0, sStem.find_last_of( '.' ) );
    const std::string sPrefix = sFolder + "/" + sStem;
    auto open = [ & ]( const std::string& sExt ) {
        FILE* p = fopen( ( sPrefix + sExt ).c_str( ), "wb" );
        if( !p )
            throw std::runtime_error( "Unable to open file " + sPrefix + sExt );
        return p;
    };
    const int64_t iFwd = (int64_t)vFwd.size( ), iRefLen = aL2[ 4 ];
    const int32_t iSaIntv = 32;
    FILE* p = open( ".bwt" ); // primary, L2[1..4], occurrence blocks
    fwrite( &iPrimary, 8, 1, p ), fwrite( aL2 + 1, 8, 4, p ), fwrite( vBwt.data( ), 4, vBwt.size( ), p ), fclose( p );
    p = open( ".sa" ); // primary, L2[1..4], interval, length, samples 1..
    fwrite( &iPrimary, 8, 1, p ), fwrite( aL2 + 1, 8, 4, p ), fwrite( &iSaIntv, 4, 1, p ), fwrite( &iRefLen, 8, 1, p );
    fwrite( vSa.data( ) + 1, 8, vSa.size( ) - 1, p ), fclose( p );
    p = open( ".pac" ); // 2 bit per base, a zero byte if the length is a multiple of 4, the length modulo 4
    fwrite( vPac.data( ), 1, (size_t)( ( iFwd + 3 ) / 4 ), p );
    if( iFwd % 4 == 0 )
        fputc( 0, p );
    fputc( (int)( iFwd % 4 ), p ), fclose( p );
    p = open( ".ann" );
    fprintf( p, "%lld %zu %d\n", (long long)iFwd, vNames.size( ), 11 );
    for( size_t i = 0; i < vNames.size( ); i++ )
        fprintf( p, "0 %s %s\n%lld %lld 0\n", vNames[ i ].c_str( ), vComments[ i ].empty( ) ? "(null)" : vComments[ i ].c_str( ),
                 (long long)vStart[ i ], (long long)vLength[ i ] );
    fclose( p );
    p = open( ".amb" );
    fprintf( p, "%lld %zu 0\n", (long long)iFwd, vNames.size( ) ), fclose( p );
    p = fopen( ( sFolder + "/" + sTitle + ".json" ).c_str( ), "wb" );
    if( !p )
        throw std::runtime_error( "Unable to open file " + sFolder + "/" + sTitle + ".json" );
    fprintf( p, "{\n    \"name\": \"%s\",\n    \"prefix\": \"%s\",\n    \"type\": \"MA Genome\",\n    \"version\": {\n"
                "        \"major\": 1,\n        \"minor\": 0\n    }\n}\n", sTitle.c_str( ), sStem.c_str( ) );
    fclose( p );
    std::cerr << "index of " << iFwd << " bases in " << vNames.size( ) << " sequences: " << sPrefix << ".*" << std::endl;
}

// -x <prefix> or, like the reference (ExecutionContext::loadGenome, execution-context.h:60-93), the genome's .json
static std::string indexPrefix( const std::string& sArg )
{
    if( sArg.size( ) < 5 || sArg.substr( sArg.size( ) - 5 ) != ".json" )
        return sArg;
    std::ifstream xIn( sArg );
    if( !xIn )
        throw std::runtime_error( "Unable to open file " + sArg );
    const std::string sText( ( std::istreambuf_iterator<char>( xIn ) ), std::istreambuf_iterator<char>( ) );
    if( sText.find( "\"MA Genome\"" ) == std::string::npos )
        throw std::runtime_error( "JSON file does not contain valid MA genome information." );
    size_t uiAt = sText.find( "\"prefix\"" );
    if( uiAt == std::string::npos )
        throw std::runtime_error( "JSON file does not contain valid MA genome information." );
    uiAt = sText.find( '"', sText.find( ':', uiAt ) );
    const size_t uiEnd = sText.find( '"', uiAt + 1 );
    const size_t uiSlash = sArg.find_last_of( '/' );
    return ( uiSlash == std::string::npos ? std::string( ) : sArg.substr( 0, uiSlash + 1 ) ) +
           sText.substr( uiAt + 1, uiEnd - uiAt - 1 );
}

int main( int argc, char** argv )
{
    std::string sIndex, sOut, sPreset = "Default";
    std::vector<std::string> vIn, vMate;
    std::vector<int> vDevices( 1, 0 );
    size_t uiBatch = 500000;
    uint32_t uiSrand = 0;
    bool bInterleaved = false, bVerbose = false, bInversions = false;
    bool bOutputM = true, bSoftClip = false, bNoSecondary = false, bNoSupplementary = false;
    int iZDropInversion = 100;
    size_t uiThreads = (size_t)std::max( 1, (int)std::thread::hardware_concurrency( ) - 3 ); // reader, device, writer
    try
    {
        for( int i = 1; i < argc; i++ )
        {
            const std::string sOpt = argv[ i ], sLow = lower( sOpt );
            auto value = [ & ]( ) -> std::string {
                if( i + 1 >= argc )
                    throw std::runtime_error( "missing value for " + sOpt );
                return argv[ ++i ];
            };
            if( sOpt == "-x" || sLow == "--index" )
                sIndex = indexPrefix( value( ) );
            else if( sOpt == "-X" || sLow == "--create_index" || sLow == "--createindex" )
            {
                const auto vParts = splitList( value( ) );
                if( vParts.size( ) != 3 )
                    throw std::runtime_error( "--Index needs exactly three parameters" );
                createIndex( vParts[ 0 ], vParts[ 1 ], vParts[ 2 ], vDevices[ 0 ] );
                return 0;
            }
            else if( sOpt == "-i" || sLow == "--in" )
                vIn = splitList( value( ) );
            else if( sOpt == "-m" || sLow == "--matein" )
                vMate = splitList( value( ) );
            else if( sOpt == "-o" || sLow == "--out" )
                sOut = value( );
            else if( sOpt == "-p" || sLow == "--presetting" )
                sPreset = value( );
            else if( sOpt == "-t" )
                uiThreads = (size_t)std::max( 1, atoi( value( ).c_str( ) ) );
            else if( sLow == "--verbose" )
                bVerbose = true;
            else if( sLow == "--interleaved" )
                bInterleaved = true;
            else if( sLow == "--detect_small_inversions" )
            { // a flag of the reference ("Detect Small Inversions"); an explicit true / false may follow
                bInversions = true;
                if( i + 1 < argc && ( lower( argv[ i + 1 ] ) == "true" || lower( argv[ i + 1 ] ) == "false" ) )
                    bInversions = lower( argv[ ++i ] ) == "true";
            }
            else if( sLow == "--z_drop_inversions" )
                iZDropInversion = atoi( value( ).c_str( ) );
            else if( sLow == "--use_m_in_cigar" || sLow == "--soft_clip" || sLow == "--omit_secondary_alignments" ||
                     sLow == "--omit_supplementary_alignments" )
            { // the writers' flags (parameter.h:726-755); an explicit true / false may follow
                bool bValue = true;
                if( i + 1 < argc && ( lower( argv[ i + 1 ] ) == "true" || lower( argv[ i + 1 ] ) == "false" ) )
                    bValue = lower( argv[ ++i ] ) == "true";
                ( sLow == "--use_m_in_cigar" ? bOutputM : sLow == "--soft_clip" ? bSoftClip
                  : sLow == "--omit_secondary_alignments" ? bNoSecondary : bNoSupplementary ) = bValue;
            }
            else if( sLow == "--device" || sLow == "--devices" )
            {
                vDevices.clear( );
                for( auto& d : splitList( value( ) ) )
                    vDevices.push_back( atoi( d.c_str( ) ) );
                if( vDevices.empty( ) )
                    throw std::runtime_error( "--Devices needs a list of CUDA devices" );
            }
            else if( sLow == "--batch" )
                uiBatch = (size_t)atoll( value( ).c_str( ) );
            else if( sLow == "--srand" )
                uiSrand = (uint32_t)atoll( value( ).c_str( ) );
            else
                throw std::runtime_error( "unknown option " + sOpt + " (this front end covers the alignment path only)" );
        }
        if( sIndex.empty( ) || vIn.empty( ) )
        {
            std::cerr << "usage: maCMD_b200 -x <index prefix> -i <reads> [-m <mates>] [-o <out.sam>] [-p <presetting>]\n";
            return argc <= 1 ? 0 : 1;
        }
        const auto tStart = std::chrono::steady_clock::now( );
        std::vector<std::unique_ptr<Aligner>> vAligners;
        for( int iDevice : vDevices )
        {
            vAligners.emplace_back( new Aligner( sIndex, sPreset, iDevice ) );
            if( !vMate.empty( ) )
                vAligners.back( )->params( ).xParams.use_paired_reads = 1;
            vAligners.back( )->params( ).bSearchInversions = bInversions;
            vAligners.back( )->params( ).iZDropInversion = iZDropInversion;
        }
        const double fStartup = std::chrono::duration<double>( std::chrono::steady_clock::now( ) - tStart ).count( );
        Aligner& xAligner = *vAligners[ 0 ];
        const bool bPaired = xAligner.params( ).xParams.use_paired_reads != 0;
        if( bPaired && vMate.empty( ) && !bInterleaved )
            throw std::runtime_error( "paired presetting: give the mates with -m (or --Interleaved)" );
        if( uiBatch < 2 )
            uiBatch = 2;
        uiBatch &= ~(size_t)1;

        FILE* pOut = sOut.empty( ) ? stdout : fopen( sOut.c_str( ), "wb" );
        if( !pOut )
            throw std::runtime_error( "Unable to open file " + sOut );
        SamWriter xWriter( xAligner.index( ).xContigs, bOutputM, bSoftClip, bNoSecondary, bNoSupplementary );
        const std::string sHead = xWriter.header( );
        fwrite( sHead.data( ), 1, sHead.size( ), pOut );

        // three host stages run concurrently: the reader parses batch k + 1 while the GPU aligns batch k and the writer
        // formats batch k - 1 on uiThreads host threads (records of different reads are independent)
        struct Batch
        {
            std::vector<NucSeq> vReads;
            size_t uiFirst = 0, uiSeq = 0;
            RawReport xRaw;
            std::vector<std::vector<Alignment>> vRecords; // with --Detect_Small_Inversions: SmallInversions' vectors
            PinnedVector<uint8_t> vSlab; // the reads as the C ABI takes them, filled by the reader's threads
            PinnedVector<int64_t> vOffsets;
            std::vector<std::string> vText; // SAM text of the batch, one piece per formatting thread
            size_t uiChunks = 0;
            bool bHasRecords = false;
        };
        const size_t uiPool = 4 + 2 * vAligners.size( );
        BoundedQueue<std::unique_ptr<Batch>> xParsed( 2 ), xAligned( 2 + vAligners.size( ) ), xFormatted( 1 ), xFree( uiPool );
        for( size_t i = 0; i < uiPool; i++ )
            xFree.push( std::make_unique<Batch>( ) );
        std::exception_ptr pReaderError, pWriterError;
        double fParse = 0, fGpu = 0, fFormat = 0, fWrite = 0;
        auto now = []( ) { return std::chrono::steady_clock::now( ); };
        auto secs = []( std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b ) {
            return std::chrono::duration<double>( b - a ).count( );
        };

        // reader, first half: the serial pass over the line ends that finds the records of a batch (the mate file on
        // its own thread); it runs ahead of the conversion of the previous batch
        struct ScanJob
        {
            std::vector<ReadStream::Record> vRecords;
            std::vector<std::shared_ptr<ReadParser>> vKeepAlive, vKeepAliveMate;
        };
        BoundedQueue<std::unique_ptr<ScanJob>> xScanned( 2 );
        std::exception_ptr pScanError;
        double fScan = 0;
        std::thread xScanner( [ & ]( ) {
            try
            {
                ReadStream xIn( vIn ), xMate( vMate );
                std::vector<ReadStream::Record> vFirst, vSecond;
                auto scan = []( ReadStream& rStream, std::vector<ReadStream::Record>& vOut, size_t uiMax,
                                std::vector<std::shared_ptr<ReadParser>>& vKeep ) {
                    vOut.clear( );
                    ReadStream::Record xR;
                    while( vOut.size( ) < uiMax && rStream.next( xR, vKeep ) )
                        vOut.push_back( xR );
                };
                bool bMore = true;
                std::vector<std::shared_ptr<ReadParser>> vLastMateKeep;
                while( bMore )
                {
                    const auto t0 = now( );
                    auto pJob = std::make_unique<ScanJob>( );
                    if( vMate.empty( ) )
                    {
                        scan( xIn, pJob->vRecords, uiBatch, pJob->vKeepAlive );
                        bMore = pJob->vRecords.size( ) == uiBatch;
                    }
                    else
                    {
                        std::exception_ptr pMateError;
                        std::thread xMateScan( [ & ]( ) {
                            try
                            {
                                scan( xMate, vSecond, uiBatch / 2, pJob->vKeepAliveMate );
                            }
                            catch( ... )
                            {
                                pMateError = std::current_exception( );
                            }
                        } );
                        try
                        {
                            scan( xIn, vFirst, uiBatch / 2, pJob->vKeepAlive );
                        }
                        catch( ... )
                        {
                            xMateScan.join( );
                            throw;
                        }
                        xMateScan.join( );
                        if( pMateError )
                            std::rethrow_exception( pMateError );
                        if( vSecond.size( ) < vFirst.size( ) )
                            throw std::runtime_error( "fewer mates than reads" );
                        if( vSecond.size( ) > vFirst.size( ) )
                            throw std::runtime_error( "more mates than reads" );
                        bMore = vFirst.size( ) == uiBatch / 2;
                        pJob->vRecords.resize( 2 * vFirst.size( ) );
                        for( size_t k = 0; k < vFirst.size( ); k++ )
                            pJob->vRecords[ 2 * k ] = vFirst[ k ], pJob->vRecords[ 2 * k + 1 ] = vSecond[ k ];
                        vLastMateKeep = pJob->vKeepAliveMate;
                    }
                    fScan += secs( t0, now( ) );
                    if( pJob->vRecords.empty( ) )
                        break;
                    if( bPaired && pJob->vRecords.size( ) % 2 )
                        throw std::runtime_error( "odd number of reads for a paired presetting" );
                    if( !xScanned.push( std::move( pJob ) ) )
                        return;
                }
                ReadStream::Record xR;
                if( !vMate.empty( ) && xMate.next( xR, vLastMateKeep ) )
                    throw std::runtime_error( "more mates than reads" );
            }
            catch( ... )
            {
                pScanError = std::current_exception( );
            }
            xScanned.close( );
        } );

        // reader, second half: the records converted to NucSeq and gathered into the batch's slab on a few threads
        std::thread xReader( [ & ]( ) {
            try
            {
                const size_t uiParseThreads = std::max<size_t>( 1, std::min<size_t>( 4, uiThreads / 3 ) );
                size_t uiDone = 0, uiSeq = 0;
                std::unique_ptr<ScanJob> pJob;
                while( xScanned.pop( pJob ) )
                {
                    std::unique_ptr<Batch> pB;
                    if( !xFree.pop( pB ) ) // recycled: the reads of a used batch keep their buffers
                        break;
                    const auto t0 = now( );
                    pB->uiFirst = uiDone, pB->uiSeq = uiSeq++;
                    const auto& vRecords = pJob->vRecords;
                    const size_t n = vRecords.size( );
                    auto& v = pB->vReads;
                    v.resize( n );
                    const size_t uiParts = std::max<size_t>( 1, std::min<size_t>( uiParseThreads, n / 4096 + 1 ) );
                    auto parallel = [ & ]( std::function<void( size_t )> fPart ) {
                        std::vector<std::exception_ptr> vErr( uiParts );
                        std::vector<std::thread> vWorkers;
                        for( size_t c = 0; c < uiParts; c++ )
                            vWorkers.emplace_back( [ &, c ]( ) {
                                try
                                {
                                    fPart( c );
                                }
                                catch( ... )
                                {
                                    vErr[ c ] = std::current_exception( );
                                }
                            } );
                        for( auto& t : vWorkers )
                            t.join( );
                        for( auto& e : vErr )
                            if( e )
                                std::rethrow_exception( e );
                    };
                    parallel( [ & ]( size_t c ) {
                        for( size_t k = n * c / uiParts; k < n * ( c + 1 ) / uiParts; k++ )
                            vRecords[ k ].pOwner->parseRecord( vRecords[ k ].uiBegin, vRecords[ k ].uiEnd, v[ k ] );
                    } );
                    // ... and gathered into the page-locked slab the device stage uploads
                    Aligner::slabOffsets( v, pB->vSlab, pB->vOffsets );
                    parallel( [ & ]( size_t c ) { Aligner::slabCopy( v, pB->vSlab, pB->vOffsets, c, uiParts ); } );
                    uiDone += n;
                    fParse += secs( t0, now( ) );
                    if( !xParsed.push( std::move( pB ) ) )
                        break;
                }
            }
            catch( ... )
            {
                pReaderError = std::current_exception( );
            }
            xScanned.close( ); // releases the scanner if this half stopped early
            xParsed.close( );
        } );

        std::thread xWriterThread( [ & ]( ) {
            try
            {
                std::unique_ptr<Batch> pNext;
                std::map<size_t, std::unique_ptr<Batch>> xPending; // batches of other devices that finished early
                size_t uiSeq = 0;
                while( xAligned.pop( pNext ) )
                {
                    xPending[ pNext->uiSeq ] = std::move( pNext );
                    while( !xPending.empty( ) && xPending.begin( )->first == uiSeq )
                    {
                    std::unique_ptr<Batch> pB = std::move( xPending.begin( )->second );
                    xPending.erase( xPending.begin( ) );
                    uiSeq++;
                    const auto t0 = now( );
                    const size_t uiUnits = pB->bHasRecords ? pB->vRecords.size( ) : pB->xRaw.units( );
                    const size_t uiChunks = std::max<size_t>( 1, std::min<size_t>( uiThreads, uiUnits / 256 + 1 ) );
                    auto& vText = pB->vText; // kept with the batch: the buffers are reused
                    vText.resize( std::max( vText.size( ), uiChunks ) );
                    for( auto& sText : vText )
                        sText.clear( );
                    pB->uiChunks = uiChunks;
                    std::vector<std::thread> vWorkers;
                    std::vector<std::exception_ptr> vErr( uiChunks );
                    for( size_t c = 0; c < uiChunks; c++ )
                        vWorkers.emplace_back( [ &, c ]( ) {
                            try
                            {
                                std::string& sText = vText[ c ];
                                std::vector<Alignment> vScratch; // reused from unit to unit
                                for( size_t u = uiUnits * c / uiChunks; u < uiUnits * ( c + 1 ) / uiChunks; u++ )
                                {
                                    if( !pB->bHasRecords )
                                        pB->xRaw.records( u, vScratch );
                                    const std::vector<Alignment>& vRec = pB->bHasRecords ? pB->vRecords[ u ] : vScratch;
                                    if( bPaired )
                                        xWriter.paired( sText, pB->vReads[ 2 * u ], pB->vReads[ 2 * u + 1 ], vRec );
                                    else
                                        xWriter.single( sText, pB->vReads[ u ], vRec );
                                }
                            }
                            catch( ... )
                            {
                                vErr[ c ] = std::current_exception( );
                            }
                        } );
                    for( auto& t : vWorkers )
                        t.join( );
                    for( auto& e : vErr )
                        if( e )
                            std::rethrow_exception( e );
                    fFormat += secs( t0, now( ) );
                    if( !xFormatted.push( std::move( pB ) ) ) // the file is written while the next batch is formatted
                        return;
                    }
                }
            }
            catch( ... )
            {
                pWriterError = std::current_exception( );
                xAligned.close( ), xFree.close( );
            }
            xFormatted.close( );
        } );

        std::exception_ptr pOutputError;
        std::thread xOutputThread( [ & ]( ) {
            try
            {
                std::unique_ptr<Batch> pB;
                size_t uiDone = 0;
                while( xFormatted.pop( pB ) )
                {
                    const auto t0 = now( );
                    for( size_t c = 0; c < pB->uiChunks; c++ )
                        if( fwrite( pB->vText[ c ].data( ), 1, pB->vText[ c ].size( ), pOut ) != pB->vText[ c ].size( ) )
                            throw std::runtime_error( "write error on " + ( sOut.empty( ) ? std::string( "stdout" ) : sOut ) );
                    fWrite += secs( t0, now( ) );
                    uiDone += pB->vReads.size( );
                    std::cerr << "\r" << uiDone << " reads aligned." << std::flush;
                    xFree.push( std::move( pB ) );
                }
            }
            catch( ... )
            {
                pOutputError = std::current_exception( );
                xFormatted.close( ), xAligned.close( ), xFree.close( );
            }
        } );

        std::vector<std::exception_ptr> vGpuError( vAligners.size( ) );
        std::vector<double> vGpuBusy( vAligners.size( ), 0.0 ), vKernelMs( vAligners.size( ), 0.0 );
        std::vector<std::thread> vGpuThreads;
        for( size_t g = 0; g < vAligners.size( ); g++ ) // one host thread per device (a context is not re-entrant)
            vGpuThreads.emplace_back( [ &, g ]( ) {
                try
                {
                    std::unique_ptr<Batch> pB;
                    while( xParsed.pop( pB ) )
                    {
                        const auto t0 = now( );
                        vAligners[ g ]->params( ).xParams.srand_base = uiSrand + (uint32_t)pB->uiFirst;
                        ma_b200_align_stats xStats;
                        memset( &xStats, 0, sizeof( xStats ) );
                        pB->vRecords.clear( );
                        pB->bHasRecords = bInversions;
                        if( bInversions ) // MappingQuality on the device, SmallInversions' DP as one more device batch,
                                          // pairing (if any) on the host
                            pB->vRecords = vAligners[ g ]->reportWithInversions( pB->vReads );
                        else
                            vAligners[ g ]->reportRaw( pB->vReads.size( ), pB->vSlab.data( ), pB->vOffsets.data( ), pB->xRaw,
                                                       &xStats );
                        vGpuBusy[ g ] += secs( t0, now( ) ), vKernelMs[ g ] += xStats.ms_total;
                        if( !xAligned.push( std::move( pB ) ) )
                            break;
                    }
                }
                catch( ... )
                {
                    vGpuError[ g ] = std::current_exception( );
                    xParsed.close( );
                }
            } );
        for( auto& t : vGpuThreads )
            t.join( );
        xParsed.close( ), xFree.close( ); // releases a reader that still waits after the devices stopped early
        double fKernels = 0;
        for( size_t g = 0; g < vGpuBusy.size( ); g++ )
            fGpu = std::max( fGpu, vGpuBusy[ g ] ), fKernels = std::max( fKernels, vKernelMs[ g ] * 1e-3 );
        xAligned.close( );
        xScanned.close( );
        xScanner.join( );
        xReader.join( );
        xWriterThread.join( );
        xOutputThread.join( );
        vGpuError.push_back( pScanError ), vGpuError.push_back( pReaderError ), vGpuError.push_back( pWriterError ),
            vGpuError.push_back( pOutputError );
        for( auto& e : vGpuError )
            if( e )
                std::rethrow_exception( e );
        if( bVerbose )
            fprintf( stderr, "\rstart-up (CUDA context, index files -> device) %.3f s, total %.3f s\n", fStartup,
                     secs( tStart, now( ) ) );
        if( bVerbose )
            fprintf( stderr, "busy seconds: scan %.3f, reader %.3f, gpu stage (upload, kernels, download; max over %zu devices) %.3f of which "
                             "kernels %.3f, format %.3f (%zu threads), write %.3f\n", fScan, fParse, vAligners.size( ), fGpu, fKernels, fFormat, uiThreads, fWrite );
        if( pOut != stdout )
            fclose( pOut );
        std::cerr << "\rdone.                         " << std::endl;
    }
    catch( std::exception& ex )
    {
        std::cerr << "Error:\n" << ex.what( ) << std::endl;
        return 1;
    }
    return 0;
}
