// SAM output of the alignment records and FASTA/FASTQ input, host C++17, header only.
//
// Restates, line for line of OUTPUT, the reference's writers with their default ("no NGMLR emulation") settings:
//   FileWriter::execute        libs/ma/src/module/fileWriter.cpp:11-156
//   PairedFileWriter::execute  libs/ma/src/module/fileWriter.cpp:158-372
//   header                     libs/ma/inc/ma/module/fileWriter.h:386-398
//   Alignment::cigarStringWithMInsteadOfXandEqual / cigarString / getSamFlag / getContig / getSamPosition /
//   getQuerySequence           libs/ma/inc/ma/container/alignment.h:367-467, 576-617
//   Pack::posInSequence / iAbsolutePosition / uiSequenceIdForPosition  libs/ma/inc/ma/container/pack.h:900-1067
// Parity: tests/golden/gold_*.sam are written by the reference's own writers (oracle/ref_dump.cpp `sam`).
// Not covered: the NGMLR tag emulation ("Emulate NGMLR's tag output", off in every preset) and the CG:B:I tag for
// CIGARs of 65 536 operations or more.
#pragma once
#include "ma_b200_modules.hpp"
#include <algorithm>
#include <cmath>
#include <fcntl.h>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#ifdef MA_B200_WITH_ZLIB // gzip-compressed input like the reference's GzFileStream (fileReader.h:286-400, WITH_ZLIB)
#include <zlib.h>
#endif

namespace libMA_b200
{

class SamWriter
{
    const ContigTable& rIdx;
    bool bOutputM, bSoftClip, bNoSecondary, bNoSupplementary;

    bool onReverse( nucSeqIndex p ) const
    {
        return p >= (nucSeqIndex)rIdx.iForwardLength;
    }
    size_t contigOf( nucSeqIndex uiPos ) const // Pack::uiSequenceIdForPosition of the position mapped to the forward strand
    {
        const int64_t a = onReverse( uiPos ) ? 2 * rIdx.iForwardLength - ( (int64_t)uiPos + 1 ) : (int64_t)uiPos;
        // the last contig that starts at or before a (fragmented assemblies have 10^5 contigs: no linear scan)
        const auto it = std::upper_bound( rIdx.vStart.begin( ), rIdx.vStart.end( ), a );
        return it == rIdx.vStart.begin( ) ? 0 : (size_t)( it - rIdx.vStart.begin( ) ) - 1;
    }
    std::string contig( const Alignment& a ) const
    {
        return rIdx.vNames[ contigOf( a.uiBeginOnRef ) ];
    }
    nucSeqIndex samPosition( const Alignment& a ) const // alignment.h:591-598, pack.h:917-920, 1063-1067
    {
        const int64_t abs = onReverse( a.uiEndOnRef ) ? 2 * rIdx.iForwardLength - ( (int64_t)a.uiEndOnRef + 1 )
                                                      : (int64_t)a.uiBeginOnRef;
        int64_t r = abs - rIdx.vStart[ contigOf( (nucSeqIndex)abs ) ];
        if( onReverse( a.uiBeginOnRef ) )
            r += 1;
        return (nucSeqIndex)( r + 1 );
    }
    uint32_t samFlag( const Alignment& a ) const
    {
        return ( onReverse( a.uiBeginOnRef ) ? 0x10u : 0u ) | ( a.bSecondary ? 0x100u : 0u ) |
               ( a.bSupplementary ? 0x800u : 0u );
    }
    std::string cigar( const Alignment& a, size_t uiQuerySize ) const
    {
        std::string s;
        cigar( s, a, uiQuerySize );
        return s;
    }
    void cigar( std::string& s, const Alignment& a, size_t uiQuerySize ) const
    {
        const bool bRev = onReverse( a.uiBeginOnRef );
        const char cClip = bSoftClip ? 'S' : 'H';
        if( bRev )
        {
            if( a.uiEndOnQuery < uiQuerySize )
                num( s, uiQuerySize - a.uiEndOnQuery ), s += cClip;
        }
        else if( a.uiBeginOnQuery > 0 )
            num( s, a.uiBeginOnQuery ), s += cClip;
        size_t uiM = 0;
        const size_t n = a.data.size( );
        for( size_t k = 0; k < n; k++ )
        {
            const auto& d = a.data[ bRev ? n - 1 - k : k ];
            if( bOutputM )
            {
                if( d.first == MatchType::insertion || d.first == MatchType::deletion )
                {
                    if( uiM > 0 )
                        num( s, uiM ), s += 'M', uiM = 0;
                    num( s, d.second ), s += d.first == MatchType::insertion ? 'I' : 'D';
                }
                else
                    uiM += d.second;
            }
            else
                num( s, d.second ), s += d.first == MatchType::missmatch   ? 'X'
                                         : d.first == MatchType::insertion ? 'I'
                                         : d.first == MatchType::deletion  ? 'D'
                                                                           : '=';
        }
        if( uiM > 0 )
            num( s, uiM ), s += 'M';
        if( bRev )
        {
            if( a.uiBeginOnQuery > 0 )
                num( s, a.uiBeginOnQuery ), s += cClip;
        }
        else if( a.uiEndOnQuery < uiQuerySize )
            num( s, uiQuerySize - a.uiEndOnQuery ), s += cClip;
    }
    static std::string text( const NucSeq& q, size_t b, size_t e, bool bComplement )
    {
        std::string s;
        if( bComplement )
        {
            s.resize( e > b ? e - b : 0 );
            for( size_t i = e, k = 0; i > b; i--, k++ )
                s[ k ] = "TGCAN"[ q.vSeq[ i - 1 ] < 4 ? q.vSeq[ i - 1 ] : 4 ];
        }
        else
        {
            e = std::min( e, q.length( ) );
            s.resize( e > b ? e - b : 0 );
            for( size_t i = b; i < e; i++ )
                s[ i - b ] = "ACGTN"[ q.vSeq[ i ] < 4 ? q.vSeq[ i ] : 4 ];
        }
        return s;
    }
    std::string segment( const Alignment& a, const NucSeq& q ) const
    {
        if( bSoftClip )
            return text( q, 0, q.length( ), onReverse( a.uiBeginOnRef ) );
        return text( q, a.uiBeginOnQuery, a.uiEndOnQuery, onReverse( a.uiBeginOnRef ) );
    }
    static std::string qual( const NucSeq& q, size_t b, size_t e ) // NucSeq::fromToQual / toQualString: never reversed
    {
        if( q.vQual.empty( ) )
            return "*";
        e = std::min( e, q.length( ) );
        return e > b ? std::string( (const char*)q.vQual.data( ) + b, e - b ) : std::string( );
    }
    static std::string mapq( const Alignment& a, bool bClamp )
    {
        if( std::isnan( a.fMappingQuality ) )
            return "255";
        const int v = static_cast<int>( std::ceil( a.fMappingQuality * 254 ) );
        return std::to_string( bClamp ? std::min( v, 255 ) : v );
    }

    static void num( std::string& s, uint64_t v )
    {
        char a[ 24 ];
        int n = 0;
        do
            a[ n++ ] = (char)( '0' + v % 10 ), v /= 10;
        while( v );
        while( n )
            s += a[ --n ];
    }
    static void text( std::string& s, const NucSeq& q, size_t b, size_t e, bool bComplement )
    {
        if( !bComplement )
            e = std::min( e, q.length( ) );
        if( e <= b )
            return;
        const size_t o = s.size( );
        s.resize( o + ( e - b ) );
        char* p = &s[ o ];
        if( bComplement )
            for( size_t i = e, k = 0; i > b; i--, k++ )
                p[ k ] = "TGCAN"[ q.vSeq[ i - 1 ] < 4 ? q.vSeq[ i - 1 ] : 4 ];
        else
            for( size_t i = b; i < e; i++ )
                p[ i - b ] = "ACGTN"[ q.vSeq[ i ] < 4 ? q.vSeq[ i ] : 4 ];
    }
    void segment( std::string& s, const Alignment& a, const NucSeq& q ) const
    {
        if( bSoftClip )
            text( s, q, 0, q.length( ), onReverse( a.uiBeginOnRef ) );
        else
            text( s, q, a.uiBeginOnQuery, a.uiEndOnQuery, onReverse( a.uiBeginOnRef ) );
    }
    static void qual( std::string& s, const NucSeq& q, size_t b, size_t e )
    {
        if( q.vQual.empty( ) )
        {
            s += '*';
            return;
        }
        e = std::min( e, q.length( ) );
        if( e > b )
            s.append( (const char*)q.vQual.data( ) + b, e - b );
    }

  public:
    // bOutputM: "Use M in CIGAR" (default true); bSoftClip: "Soft clip" (default false); parameter.h:743-755
    explicit SamWriter( const ContigTable& rIndex, bool bOutputM = true, bool bSoftClip = false, bool bNoSecondary = false,
                        bool bNoSupplementary = false )
        : rIdx( rIndex ), bOutputM( bOutputM ), bSoftClip( bSoftClip ), bNoSecondary( bNoSecondary ),
          bNoSupplementary( bNoSupplementary )
    {}
    std::string header( ) const
    {
        std::string s;
        for( size_t i = 0; i < rIdx.vNames.size( ); i++ )
            s += "@SQ\tSN:" + rIdx.vNames[ i ] + "\tLN:" + std::to_string( rIdx.vLength[ i ] ) + "\n";
        return s + "@PG\tID:ma\tPN:ma\tVN:0.1.0\tCL:na\n";
    }
    // FileWriter::execute: the records of one read (MappingQuality's result vector)
    std::string single( const NucSeq& q, const std::vector<Alignment>& v ) const
    {
        std::string s;
        single( s, q, v );
        return s;
    }
    // the same, appended to a caller-owned buffer (no temporaries: the writer threads of maCMD_b200 use these)
    void single( std::string& s, const NucSeq& q, const std::vector<Alignment>& v ) const
    {
        const size_t uiBefore = s.size( );
        for( const Alignment& a : v )
        {
            if( a.uiLength == 0 || ( bNoSecondary && a.bSecondary ) || ( bNoSupplementary && a.bSupplementary ) )
                continue;
            s.append( q.sName ), s += '\t', num( s, samFlag( a ) ), s += '\t', s.append( contig( a ) ), s += '\t';
            num( s, samPosition( a ) ), s += '\t', s.append( mapq( a, false ) ), s += '\t';
            cigar( s, a, q.length( ) ), s.append( "\t*\t0\t0\t" );
            segment( s, a, q ), s += '\t', qual( s, q, a.uiBeginOnQuery, a.uiEndOnQuery ), s += '\n';
        }
        if( v.empty( ) )
        {
            s.append( q.sName ), s.append( "\t4\t*\t0\t255\t*\t*\t0\t0\t" ), text( s, q, 0, q.length( ), false );
            s += '\t', qual( s, q, 0, q.length( ) ), s += '\n';
        }
        if( s.size( ) == uiBefore )
        {
            s.append( q.sName ), s.append( "\t4\t*\t0\t0\t*\t*\t0\t0\t" ), text( s, q, 0, q.length( ), false );
            s += '\t', qual( s, q, 0, q.length( ) ), s += '\n';
        }
    }
    // PairedFileWriter::execute: the records of one pair (PairedReads' result vector; bFirst tells the mate)
    std::string paired( const NucSeq& q1, const NucSeq& q2, const std::vector<Alignment>& v ) const
    {
        std::string s;
        paired( s, q1, q2, v );
        return s;
    }
    void paired( std::string& s, const NucSeq& q1, const NucSeq& q2, const std::vector<Alignment>& v ) const
    {
        bool bHas1 = false, bHas2 = false;
        // PairedReads links the two chosen alignments (xStats.pOther); a passed-through vector has no links
        const bool bLinked = v.size( ) == 2 && v[ 0 ].bFirst != v[ 1 ].bFirst;
        for( size_t k = 0; k < v.size( ); k++ )
        {
            const Alignment& a = v[ k ];
            if( a.uiLength == 0 || ( bNoSecondary && a.bSecondary ) || ( bNoSupplementary && a.bSupplementary ) )
                continue;
            ( a.bFirst ? bHas1 : bHas2 ) = true;
            const NucSeq& q = a.bFirst ? q1 : q2;
            uint32_t flag = samFlag( a ) | 0x1u | 0x2u | ( a.bFirst ? 0x40u : 0x80u );
            const size_t uiRef = contigOf( a.uiBeginOnRef );
            size_t uiRefOther = uiRef;
            nucSeqIndex uiPosOther = 0;
            if( bLinked )
            {
                const Alignment& o = v[ 1 - k ];
                if( onReverse( o.uiBeginOnRef ) )
                    flag |= 0x20u;
                uiRefOther = contigOf( o.uiBeginOnRef );
                uiPosOther = samPosition( o );
            }
            // (the CIGAR's clip lengths use the FIRST mate's length for both mates, fileWriter.cpp:196-198)
            s.append( q.sName ), s += '\t', num( s, flag ), s += '\t', s.append( rIdx.vNames[ uiRef ] ), s += '\t';
            num( s, samPosition( a ) ), s += '\t', s.append( mapq( a, true ) ), s += '\t';
            cigar( s, a, q1.length( ) ), s += '\t';
            if( !bLinked )
                s.append( "*\t0" );
            else
            { // the other contig by NAME: "=" if the names are equal
                if( rIdx.vNames[ uiRefOther ] == rIdx.vNames[ uiRef ] )
                    s += '=';
                else
                    s.append( rIdx.vNames[ uiRefOther ] );
                s += '\t', num( s, uiPosOther );
            }
            s.append( "\t0\t" ), segment( s, a, q ), s += '\t', qual( s, q, a.uiBeginOnQuery, a.uiEndOnQuery ), s += '\n';
        }
        if( !bHas1 && !bHas2 )
        {
            s.append( q1.sName ), s += '\t', num( s, 0x4 | 0x1 | 0x40 | 0x8 ), s.append( "\t*\t0\t0\t*\t*\t0\t0\t" );
            text( s, q1, 0, q1.length( ), false ), s += '\t', qual( s, q1, 0, q1.length( ) ), s += '\n';
            s.append( q2.sName ), s += '\t', num( s, 0x4 | 0x1 | 0x80 | 0x8 ), s.append( "\t*\t0\t0\t*\t*\t0\t0\t" );
            text( s, q2, 0, q2.length( ), false ), s += '\t', qual( s, q2, 0, q2.length( ) ), s += '\n';
        }
        else if( !bHas1 || !bHas2 )
        {
            const Alignment& a0 = v[ 0 ];
            const NucSeq& q = !bHas1 ? q1 : q2;
            s.append( q.sName ), s += '\t', num( s, 0x4 | 0x1 | ( !bHas1 ? 0x40 : 0x80 ) ), s += '\t';
            s.append( contig( a0 ) ), s += '\t', num( s, samPosition( a0 ) ), s.append( "\t0\t*\t=\t" );
            num( s, samPosition( a0 ) ), s.append( "\t0\t" ), text( s, q, 0, q.length( ), false ), s.append( "\t*\n" );
        }
    }
};

// FileReader::execute (libs/ma/src/module/fileReader.cpp:37-203, default build: WITH_QUALITY == 1): (multi-)FASTA and
// FASTQ with multi-line records, LF / CR / CRLF line ends (fileReader.h:151-186), the name ends at the first blank,
// trailing characters of a sequence line that are no IUPAC nucleotide codes are cut (fileReader.cpp:12-27), every
// letter other than ACGTacgt becomes N (nucSeq.cpp:17-28).
class ReadParser
{
    // the file, memory mapped (falls back to reading it when it cannot be mapped, e.g. a pipe)
    const char* pData = nullptr;
    size_t uiSize = 0, uiPos = 0;
    void* pMap = nullptr;
    std::string sBuffer;

    // one line [b, e) starting at uiAt; uiAt moves behind its LF / CR / CR LF
    static void line( const char* p, size_t uiEnd, size_t& uiAt, size_t& b, size_t& e )
    {
        b = uiAt;
        const char* pLf = (const char*)memchr( p + uiAt, '\n', uiEnd - uiAt );
        const size_t uiLf = pLf ? (size_t)( pLf - p ) : uiEnd;
        const char* pCr = (const char*)memchr( p + uiAt, '\r', uiLf - uiAt );
        if( pCr )
        {
            e = (size_t)( pCr - p );
            uiAt = e + 1;
            if( uiAt < uiEnd && p[ uiAt ] == '\n' )
                uiAt++;
        }
        else
        {
            e = uiLf;
            uiAt = uiLf < uiEnd ? uiLf + 1 : uiEnd;
        }
    }
    static bool validNuc( char c )
    {
        switch( c )
        {
            case 'A': case 'C': case 'G': case 'T': case 'N': case 'U': case 'R': case 'Y': case 'K': case 'M': case 'S':
            case 'W': case 'B': case 'D': case 'H': case 'V':
            case 'a': case 'c': case 'g': case 't': case 'n': case 'u': case 'r': case 'y': case 'k': case 'm': case 's':
            case 'w': case 'b': case 'd': case 'h': case 'v':
                return true;
            default:
                return false;
        }
    }
    static void append( NucSeq& q, const char* p, size_t b, size_t e )
    {
        while( e > b && !validNuc( p[ e - 1 ] ) ) // fileReader.cpp:12-27
            e--;
        const size_t o = q.vSeq.size( );
        q.vSeq.resize( o + ( e - b ) );
        static const struct Lut
        {
            uint8_t a[ 256 ];
            Lut( )
            {
                memset( a, 4, sizeof( a ) ); // every other letter becomes N (nucSeq.cpp:17-28)
                a[ 'A' ] = a[ 'a' ] = 0, a[ 'C' ] = a[ 'c' ] = 1, a[ 'G' ] = a[ 'g' ] = 2, a[ 'T' ] = a[ 't' ] = 3;
            }
        } xLut;
        uint8_t* pOut = q.vSeq.data( ) + o;
        for( size_t i = b; i < e; i++ )
            pOut[ i - b ] = xLut.a[ (unsigned char)p[ i ] ];
    }
    static void name( NucSeq& q, const char* p, size_t b, size_t e )
    {
        const char* pBlank = (const char*)memchr( p + b, ' ', e - b );
        q.sName.assign( p + b + 1, ( pBlank ? (size_t)( pBlank - p ) : e ) - b - 1 );
    }
    // FileReader::execute on [uiAt, uiEnd): with STORE the read goes to *pQ, without it only uiAt moves to the record
    // that follows (same control flow, so that record boundaries found by one pass hold for the other)
    template <bool STORE> static void parseOne( const char* p, const size_t uiEnd, size_t& uiAt, NucSeq* pQ )
    {
        size_t b, e;
        if( STORE ) // cleared, not replaced: a recycled NucSeq keeps its buffers
            pQ->vSeq.clear( ), pQ->vQual.clear( ), pQ->sName.clear( );
        if( p[ uiAt ] == '>' )
        {
            line( p, uiEnd, uiAt, b, e );
            if( STORE )
                name( *pQ, p, b, e );
            while( uiAt < uiEnd && p[ uiAt ] != '>' && p[ uiAt ] != ' ' )
            {
                line( p, uiEnd, uiAt, b, e );
                if( STORE && e > b )
                    append( *pQ, p, b, e );
            }
        }
        else if( p[ uiAt ] == '@' )
        {
            line( p, uiEnd, uiAt, b, e );
            if( STORE )
                name( *pQ, p, b, e );
            while( uiAt < uiEnd && p[ uiAt ] != '+' && p[ uiAt ] != ' ' )
            {
                line( p, uiEnd, uiAt, b, e );
                if( STORE && e > b )
                    append( *pQ, p, b, e );
            }
            if( STORE )
                pQ->vQual.assign( pQ->vSeq.size( ), 126 ); // NucSeq::addQuality fills with 126 (nucSeq.h resize default)
            line( p, uiEnd, uiAt, b, e );
            if( e > b && p[ b ] == '+' )
            {
                size_t uiQ = 0;
                while( uiAt < uiEnd && ( p[ uiAt ] != '@' || uiQ == 0 ) )
                {
                    line( p, uiEnd, uiAt, b, e );
                    const size_t n = e - b;
                    if( n == 0 )
                        continue;
                    if( STORE )
                    {
                        if( uiQ + n > pQ->vQual.size( ) )
                            pQ->vQual.resize( uiQ + n, 126 );
                        memcpy( pQ->vQual.data( ) + uiQ, p + b, n );
                    }
                    uiQ += n;
                }
            }
        }
        else
            throw std::runtime_error( "Error while reading file.\nIs your input really in FASTA/Q format?" );
        if( STORE && pQ->length( ) == 0 )
            throw std::runtime_error( "found empty read: " + pQ->sName );
        while( uiAt < uiEnd && p[ uiAt ] != '>' && p[ uiAt ] != '@' ) // advanceTillNext
            uiAt++;
    }

  public:
    explicit ReadParser( const std::string& sFileName )
    {
        const int fd = open( sFileName.c_str( ), O_RDONLY );
        if( fd < 0 )
            throw std::runtime_error( "Unable to open file " + sFileName );
        unsigned char aMagic[ 2 ] = { 0, 0 };
        const bool bGzip = pread( fd, aMagic, 2, 0 ) == 2 && aMagic[ 0 ] == 0x1f && aMagic[ 1 ] == 0x8b;
        struct stat xStat;
        if( bGzip )
        { // inflated into memory as a whole (the record scan needs random access)
            close( fd );
#ifdef MA_B200_WITH_ZLIB
            gzFile pGz = gzopen( sFileName.c_str( ), "rb" );
            if( !pGz )
                throw std::runtime_error( "Unable to open file " + sFileName );
            gzbuffer( pGz, 1 << 20 );
            std::vector<char> vChunk( 1 << 22 );
            int n;
            while( ( n = gzread( pGz, vChunk.data( ), (unsigned)vChunk.size( ) ) ) > 0 )
                sBuffer.append( vChunk.data( ), (size_t)n );
            const bool bBad = n < 0;
            gzclose( pGz );
            if( bBad )
                throw std::runtime_error( "Error while inflating " + sFileName );
            pData = sBuffer.data( ), uiSize = sBuffer.size( );
            return;
#else
            throw std::runtime_error( sFileName + " is gzip-compressed: build with -DMA_B200_WITH_ZLIB -lz" );
#endif
        }
        if( fstat( fd, &xStat ) == 0 && S_ISREG( xStat.st_mode ) && xStat.st_size > 0 )
        {
            void* pM = mmap( nullptr, (size_t)xStat.st_size, PROT_READ, MAP_PRIVATE, fd, 0 );
            if( pM != MAP_FAILED )
            {
                pMap = pM, pData = (const char*)pM, uiSize = (size_t)xStat.st_size;
                madvise( pM, uiSize, MADV_SEQUENTIAL );
            }
        }
        if( !pMap )
        {
            char aBuf[ 1 << 16 ];
            ssize_t n;
            while( ( n = read( fd, aBuf, sizeof( aBuf ) ) ) > 0 )
                sBuffer.append( aBuf, (size_t)n );
            pData = sBuffer.data( ), uiSize = sBuffer.size( );
        }
        close( fd );
    }
    ReadParser( const ReadParser& ) = delete;
    ReadParser& operator=( const ReadParser& ) = delete;
    ~ReadParser( )
    {
        if( pMap )
            munmap( pMap, uiSize );
    }
    // next read; false at the end of the file
    bool next( NucSeq& q )
    {
        if( uiPos >= uiSize )
            return false;
        parseOne<true>( pData, uiSize, uiPos, &q );
        return true;
    }
    // Two-pass use for host threads: nextRecord() is the serial pass (line ends and first characters only) that yields
    // the byte range of the next record; parseRecord() converts one such range and may run on any thread.
    bool nextRecord( size_t& uiBegin, size_t& uiEnd )
    {
        if( uiPos >= uiSize )
            return false;
        uiBegin = uiPos;
        parseOne<false>( pData, uiSize, uiPos, nullptr );
        uiEnd = uiPos;
        return true;
    }
    void parseRecord( size_t uiBegin, size_t uiEnd, NucSeq& q ) const
    {
        parseOne<true>( pData, uiEnd, uiBegin, &q );
    }
};

} // namespace libMA_b200
