// CPU test of include/ma_b200_sam.hpp: reads the reported alignments of every read / pair as text (produced from the
// golden dumps of the compiled reference by tests/test_sam_writer.py) and prints the SAM file.
//   test_sam <index prefix> <paired 0|1> [<use M 0|1> <soft clip 0|1> <omit secondary 0|1> <omit supplementary 0|1>] < records
// records:  Q <name> <sequence>      one per read, in read order
//           A <unit> <first 0|1> <begin_q> <end_q> <begin_ref> <end_ref> <score> <length> <sec> <supp> <mapq bits> <type:len>...
//           (unit = read index, or pair index for paired output; records of a unit in result-vector order)
#include "../../include/ma_b200_sam.hpp"
#include <iostream>
#include <map>

using namespace libMA_b200;

int main( int argc, char** argv )
{
    if( argc < 3 )
        return 2;
    ContigTable xC;
    xC.vLoad( argv[ 1 ] );
    const bool bPaired = atoi( argv[ 2 ] ) != 0;
    std::vector<NucSeq> vQ;
    std::map<size_t, std::vector<Alignment>> xUnits;
    std::string sLine;
    while( std::getline( std::cin, sLine ) )
    {
        std::istringstream in( sLine );
        std::string sKind;
        in >> sKind;
        if( sKind == "Q" )
        {
            std::string sName, sSeq;
            in >> sName >> sSeq;
            vQ.emplace_back( sSeq == "-" ? std::string( ) : sSeq );
            vQ.back( ).sName = sName;
        }
        else if( sKind == "A" )
        {
            size_t uiUnit;
            int iFirst, iSec, iSupp;
            long long iBits;
            Alignment a;
            in >> uiUnit >> iFirst >> a.uiBeginOnQuery >> a.uiEndOnQuery >> a.uiBeginOnRef >> a.uiEndOnRef >> a.iScore >>
                a.uiLength >> iSec >> iSupp >> iBits;
            a.bFirst = iFirst != 0, a.bSecondary = iSec != 0, a.bSupplementary = iSupp != 0;
            memcpy( &a.fMappingQuality, &iBits, 8 );
            std::string sRun;
            while( in >> sRun )
                a.data.emplace_back( (MatchType)atoi( sRun.c_str( ) ), strtoull( sRun.c_str( ) + sRun.find( ':' ) + 1, nullptr, 10 ) );
            xUnits[ uiUnit ].push_back( a );
        }
    }
    auto flag = [ & ]( int i, bool bDefault ) { return argc > i ? atoi( argv[ i ] ) != 0 : bDefault; };
    SamWriter xW( xC, flag( 3, true ), flag( 4, false ), flag( 5, false ), flag( 6, false ) );
    std::string sOut = xW.header( );
    if( bPaired )
        for( size_t p = 0; 2 * p + 1 < vQ.size( ); p++ )
            sOut += xW.paired( vQ[ 2 * p ], vQ[ 2 * p + 1 ], xUnits[ p ] );
    else
        for( size_t i = 0; i < vQ.size( ); i++ )
            sOut += xW.single( vQ[ i ], xUnits[ i ] );
    fwrite( sOut.data( ), 1, sOut.size( ), stdout );
    return 0;
}
