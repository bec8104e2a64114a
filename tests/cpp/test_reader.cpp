// CPU test of ReadParser (include/ma_b200_sam.hpp): prints name, sequence, quality of every read of a FASTA/FASTQ file
// in the format of `ref_dump reads` (the reference's own FileReader).
#include "../../include/ma_b200_sam.hpp"
#include <iostream>
using namespace libMA_b200;
int main( int argc, char** argv )
{
    if( argc < 2 )
        return 2;
    ReadParser xP( argv[ 1 ] );
    NucSeq q;
    while( xP.next( q ) )
    {
        std::string s, ql( q.vQual.begin( ), q.vQual.end( ) );
        for( auto c : q.vSeq )
            s += "ACGTN"[ c < 4 ? c : 4 ];
        std::cout << q.sName << "\t" << s << "\t" << ( q.vQual.empty( ) ? std::string( "*" ) : ql ) << "\n";
    }
    return 0;
}
