// CPU test of ReadParser's two-pass interface (include/ma_b200_sam.hpp): nextRecord() + parseRecord() on worker
// threads must deliver the reads next() delivers; prints them in the format of test_reader.cpp.
#include "../../include/ma_b200_sam.hpp"
#include <iostream>
#include <thread>
using namespace libMA_b200;
int main( int argc, char** argv )
{
    if( argc < 2 )
        return 2;
    try
    {
        ReadParser xScan( argv[ 1 ] );
        std::vector<std::pair<size_t, size_t>> vRanges;
        size_t b, e;
        while( xScan.nextRecord( b, e ) )
            vRanges.emplace_back( b, e );
        std::vector<NucSeq> vReads( vRanges.size( ) );
        std::vector<std::thread> vThreads;
        std::exception_ptr aError[ 3 ];
        for( int t = 0; t < 3; t++ )
            vThreads.emplace_back( [ &, t ]( ) {
                try
                {
                    for( size_t i = t; i < vRanges.size( ); i += 3 )
                        xScan.parseRecord( vRanges[ i ].first, vRanges[ i ].second, vReads[ i ] );
                }
                catch( ... )
                {
                    aError[ t ] = std::current_exception( );
                }
            } );
        for( auto& t : vThreads )
            t.join( );
        for( auto& e : aError )
            if( e )
                std::rethrow_exception( e );
        // a recycled NucSeq must not keep anything of the read it held before
        ReadParser xSerial( argv[ 1 ] );
        NucSeq q;
        q.vSeq.assign( 1000, 9 ), q.vQual.assign( 1000, 9 ), q.sName = "stale";
        size_t i = 0;
        while( xSerial.next( q ) )
        {
            if( i >= vReads.size( ) || q.vSeq != vReads[ i ].vSeq || q.vQual != vReads[ i ].vQual || q.sName != vReads[ i ].sName )
            {
                std::cerr << "two-pass result differs at read " << i << std::endl;
                return 1;
            }
            i++;
        }
        if( i != vReads.size( ) )
            return 1;
        for( auto& r : vReads )
        {
            std::string s, ql( r.vQual.begin( ), r.vQual.end( ) );
            for( auto c : r.vSeq )
                s += "ACGTN"[ c < 4 ? c : 4 ];
            std::cout << r.sName << "\t" << s << "\t" << ( r.vQual.empty( ) ? std::string( "*" ) : ql ) << "\n";
        }
    }
    catch( std::exception& ex )
    {
        std::cerr << "Error: " << ex.what( ) << std::endl;
        return 3;
    }
    return 0;
}
