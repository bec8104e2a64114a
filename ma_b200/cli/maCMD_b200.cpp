// maCMD_b200 — command-line front end of the B200 alignment path with the reference CLI's options for that path
// (cmdMa.cpp:252-431): reads FASTA/FASTQ like FileReader, aligns batches of reads on the GPU through
// the C ABI (libma_b200.so) and writes the SAM text of the reference's FileWriter / PairedFileWriter.
//
//   maCMD_b200 -x <index prefix> -i <reads.fq[,more.fq]> [-m <mates.fq[,more.fq]>] [-o <out.sam>] [-p <presetting>]
//
//   -x, --Index        prefix of the reference's index files (.bwt .sa .pac .ann .amb), as written by maCMD --Create_Index
//   -i, --In           FASTA / FASTQ file(s), comma separated
//   -m, --MateIn       mate file(s); switches "Use Paired Reads" on like the reference (cmdMa.cpp:323-330)
//   -o, --Out          SAM file (default: standard output)
//   -p, --Presetting   Default | Illumina | Illumina_Paired | PacBio | Nanopore (default: Default)
//   -t                 accepted and ignored (the reference's host thread count)
//   --Interleaved      with a paired presetting and no -m: reads 2k, 2k+1 of -i are mates (not in the reference)
//   --Device <n>       CUDA device (default 0)
//   --Batch <n>        reads per GPU batch (default 1000000; pairs are never split)
//   --Srand <n>        RANSAC stream of read i is srand(n + i) (parity contract; default 0)
//
// Records are written in input order. There is no CPU path: the program fails when no CUDA device is present.
#include "../../include/ma_b200_modules.hpp"
#include "../../include/ma_b200_sam.hpp"
#include <cstdio>
#include <iostream>

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

// one stream of reads over several files
class ReadStream
{
    std::vector<std::string> vFiles;
    size_t uiFile = 0;
    std::unique_ptr<ReadParser> pParser;

  public:
    explicit ReadStream( std::vector<std::string> v ) : vFiles( std::move( v ) )
    {}
    bool next( NucSeq& q )
    {
        while( true )
        {
            if( !pParser )
            {
                if( uiFile >= vFiles.size( ) )
                    return false;
                pParser.reset( new ReadParser( vFiles[ uiFile++ ] ) );
            }
            if( pParser->next( q ) )
                return true;
            pParser.reset( );
        }
    }
};

int main( int argc, char** argv )
{
    std::string sIndex, sOut, sPreset = "Default";
    std::vector<std::string> vIn, vMate;
    int iDevice = 0;
    size_t uiBatch = 1000000;
    uint32_t uiSrand = 0;
    bool bInterleaved = false;
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
                sIndex = value( );
            else if( sOpt == "-i" || sLow == "--in" )
                vIn = splitList( value( ) );
            else if( sOpt == "-m" || sLow == "--matein" )
                vMate = splitList( value( ) );
            else if( sOpt == "-o" || sLow == "--out" )
                sOut = value( );
            else if( sOpt == "-p" || sLow == "--presetting" )
                sPreset = value( );
            else if( sOpt == "-t" )
                value( );
            else if( sLow == "--interleaved" )
                bInterleaved = true;
            else if( sLow == "--device" )
                iDevice = atoi( value( ).c_str( ) );
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
        Aligner xAligner( sIndex, sPreset, iDevice );
        if( !vMate.empty( ) )
            xAligner.params( ).xParams.use_paired_reads = 1;
        const bool bPaired = xAligner.params( ).xParams.use_paired_reads != 0;
        if( bPaired && vMate.empty( ) && !bInterleaved )
            throw std::runtime_error( "paired presetting: give the mates with -m (or --Interleaved)" );
        if( uiBatch < 2 )
            uiBatch = 2;
        uiBatch &= ~(size_t)1;

        FILE* pOut = sOut.empty( ) ? stdout : fopen( sOut.c_str( ), "wb" );
        if( !pOut )
            throw std::runtime_error( "Unable to open file " + sOut );
        SamWriter xWriter( xAligner.index( ).xContigs );
        const std::string sHead = xWriter.header( );
        fwrite( sHead.data( ), 1, sHead.size( ), pOut );

        ReadStream xIn( vIn ), xMate( vMate );
        std::vector<NucSeq> vReads;
        size_t uiDone = 0;
        bool bMore = true;
        while( bMore )
        {
            vReads.clear( );
            NucSeq xQ;
            while( vReads.size( ) < uiBatch && ( bMore = xIn.next( xQ ) ) )
            {
                vReads.push_back( xQ );
                if( !vMate.empty( ) )
                {
                    if( !xMate.next( xQ ) )
                        throw std::runtime_error( "fewer mates than reads" );
                    vReads.push_back( xQ );
                }
            }
            if( vReads.empty( ) )
                break;
            if( bPaired && vReads.size( ) % 2 )
                throw std::runtime_error( "odd number of reads for a paired presetting" );
            xAligner.params( ).xParams.srand_base = uiSrand + (uint32_t)uiDone;
            auto vRep = xAligner.report( vReads );
            std::string sText;
            if( bPaired )
                for( size_t p = 0; p < vRep.size( ); p++ )
                    sText += xWriter.paired( vReads[ 2 * p ], vReads[ 2 * p + 1 ], vRep[ p ] );
            else
                for( size_t k = 0; k < vRep.size( ); k++ )
                    sText += xWriter.single( vReads[ k ], vRep[ k ] );
            fwrite( sText.data( ), 1, sText.size( ), pOut );
            uiDone += vReads.size( );
            std::cerr << "\r" << uiDone << " reads aligned." << std::flush;
        }
        if( !vMate.empty( ) && xMate.next( vReads.emplace_back( ) ) )
            throw std::runtime_error( "more mates than reads" );
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
