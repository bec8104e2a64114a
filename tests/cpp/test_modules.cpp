// C++ host-side mirror of the reference's module interface (include/ma_b200_modules.hpp) driven like the
// reference's own probe: load index, select preset, run the modules over a read file.
//   test_modules <index prefix> <reads.txt> <preset> <srand_base>   -> prints one line per alignment
#include "../../include/ma_b200_modules.hpp"
#include "../../include/ma_b200_sam.hpp"
#include <cstdio>
#include <iostream>

using namespace libMA_b200;

int main( int argc, char** argv )
{
    if( argc < 5 )
        return 2;
    try
    {
        Aligner xAligner( argv[ 1 ], argv[ 3 ] );
        xAligner.params( ).xParams.srand_base = (uint32_t)atoll( argv[ 4 ] );
        std::vector<NucSeq> vReads;
        std::ifstream in( argv[ 2 ] );
        bool bNamed = false;
        if( in.peek( ) == '>' || in.peek( ) == '@' )
        { // FASTA / FASTQ through the reader that mirrors the reference's FileReader
            ReadParser xParser( argv[ 2 ] );
            NucSeq xQ;
            while( xParser.next( xQ ) )
                vReads.push_back( xQ );
            bNamed = true;
        }
        std::string line;
        while( !bNamed && std::getline( in, line ) )
            if( !line.empty( ) )
                vReads.emplace_back( line );
        ma_b200_align_stats st;
        if( argc > 5 && std::string( argv[ 5 ] ) == "report" )
        { // what the writer receives: MappingQuality per read / PairedReads per pair (paired presets)
            auto vRep = xAligner.report( vReads, &st );
            for( size_t i = 0; i < vRep.size( ); i++ )
                for( auto& a : vRep[ i ] )
                {
                    long long bits;
                    memcpy( &bits, &a.fMappingQuality, 8 );
                    printf( "%zu %d %llu %llu %llu %llu %lld %d %lld\n", i, a.bFirst ? 0 : 1,
                            (unsigned long long)a.uiBeginOnQuery, (unsigned long long)a.uiEndOnQuery,
                            (unsigned long long)a.uiBeginOnRef, (unsigned long long)a.uiEndOnRef, (long long)a.score( ),
                            (int)a.bSecondary | ( (int)a.bSupplementary << 1 ), bits );
                }
            return 0;
        }
        if( argc > 5 && std::string( argv[ 5 ] ) == "sam" )
        { // SAM text as the reference's FileWriter / PairedFileWriter write it (reads named r<i>)
            for( size_t i = 0; !bNamed && i < vReads.size( ); i++ )
                vReads[ i ].sName = "r" + std::to_string( i );
            auto vRep = xAligner.report( vReads, &st );
            SamWriter xW( xAligner.index( ).xContigs );
            std::string sOut = xW.header( );
            if( xAligner.params( ).xParams.use_paired_reads )
                for( size_t p = 0; p < vRep.size( ); p++ )
                    sOut += xW.paired( vReads[ 2 * p ], vReads[ 2 * p + 1 ], vRep[ p ] );
            else
                for( size_t i = 0; i < vRep.size( ); i++ )
                    sOut += xW.single( vReads[ i ], vRep[ i ] );
            fwrite( sOut.data( ), 1, sOut.size( ), stdout );
            return 0;
        }
        auto vAln = xAligner.align( vReads, &st );
        for( size_t i = 0; i < vAln.size( ); i++ )
            for( auto& a : vAln[ i ] )
            {
                printf( "%zu %llu %llu %llu %llu %lld %u %llu", i, (unsigned long long)a.uiBeginOnQuery,
                        (unsigned long long)a.uiEndOnQuery, (unsigned long long)a.uiBeginOnRef,
                        (unsigned long long)a.uiEndOnRef, (long long)a.score( ), a.index_of_strip,
                        (unsigned long long)a.uiLength );
                for( auto& d : a.data )
                    printf( " %d:%llu", (int)d.first, (unsigned long long)d.second );
                printf( "\n" );
            }
        // module-by-module use, like BinarySeeding(params).execute(fm_index, query) in setupaligner.py:24-44
        ParameterSetManager xP;
        xP.setSelected( argv[ 3 ] );
        FMIndex xIdx;
        xIdx.vLoad( argv[ 1 ] );
        auto vSeg = BinarySeeding( xP ).execute( xIdx, vReads );
        auto vSeeds = BinarySeeding( xP ).seed( xIdx, vReads );
        size_t nSeg = 0, nSeeds = 0;
        for( auto& v : vSeg )
            nSeg += v.size( );
        for( auto& v : vSeeds )
            nSeeds += v.vContent.size( );
        fprintf( stderr, "reads %zu segments %zu seeds %zu launches %d\n", vReads.size( ), nSeg, nSeeds, st.launches );
    }
    catch( const std::exception& e )
    {
        std::cerr << "exception: " << e.what( ) << std::endl;
        return 3;
    }
    return 0;
}
