// pybind11 bindings of the host-side module mirror (include/ma_b200_modules.hpp) under the names the reference
// registers for the path in libs/ma/src/util/export.cpp:38-67 and in the containers' export functions
// (alignment.cpp:328-364, seed.cpp:110-146, nucSeq.cpp, fMIndex.cpp): ParameterSetManager, NucSeq, Seed, Seeds,
// Segment, Alignment, FMIndex, BinarySeeding, Harmonization, NeedlemanWunsch, MappingQuality, PairedReads.
// The reference's own pybind module does not build against Python 3.12; this one is built against
// the system pybind11 by __graft_entry__.build() / `make -C ma_b200/pybind`. Modules take a batch (list of NucSeq)
// where the reference takes one read — one GPU launch per read would waste the device.
#include "../../include/ma_b200_sam.hpp"
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

namespace py = pybind11;
using namespace libMA_b200;

PYBIND11_MODULE( ma_b200_py, m )
{
    m.doc( ) = "B200-native hot path of MA behind the reference's module names (batched)";

    py::class_<ParameterSetManager>( m, "ParameterSetManager" )
        .def( py::init<>( ) )
        .def( "set_selected", &ParameterSetManager::setSelected )
        .def( "by_name", // global / selected parameters by the flattened names of ma_b200_params
              []( ParameterSetManager& p, const std::string& n ) -> double {
                  const ma_b200_params& x = p.xParams;
                  if( n == "match" ) return x.match;
                  if( n == "mismatch" ) return x.mismatch;
                  if( n == "max_ambiguity" ) return x.max_ambiguity;
                  if( n == "seeding_technique" ) return x.seeding_technique;
                  if( n == "use_paired_reads" ) return x.use_paired_reads;
                  if( n == "min_alignment_score" ) return x.min_alignment_score;
                  throw std::runtime_error( "unknown parameter " + n );
              } )
        .def_property( "srand_base", []( ParameterSetManager& p ) { return p.xParams.srand_base; },
                       []( ParameterSetManager& p, uint32_t v ) { p.xParams.srand_base = v; } )
        .def_readwrite( "detect_small_inversions", &ParameterSetManager::bSearchInversions )
        .def_readwrite( "z_drop_inversions", &ParameterSetManager::iZDropInversion )
        .def_property( "use_paired_reads", []( ParameterSetManager& p ) { return p.xParams.use_paired_reads != 0; },
                       []( ParameterSetManager& p, bool v ) { p.xParams.use_paired_reads = v ? 1 : 0; } );

    py::class_<NucSeq>( m, "NucSeq" )
        .def( py::init<>( ) )
        .def( py::init<const std::string&>( ) )
        .def_readwrite( "name", &NucSeq::sName )
        .def( "__len__", &NucSeq::length )
        .def( "__getitem__", []( const NucSeq& s, size_t i ) { return (int)s.vSeq.at( i ); } )
        .def( "__str__", []( const NucSeq& s ) {
            std::string r;
            for( auto c : s.vSeq )
                r += "ACGTN"[ c < 4 ? c : 4 ];
            return r;
        } );

    py::class_<Segment>( m, "Segment" )
        .def_readonly( "start", &Segment::uiStart )
        .def_readonly( "size", &Segment::uiSize )
        .def_readonly( "sa_start", &Segment::iSaStart )
        .def_readonly( "sa_start_rev_comp", &Segment::iSaStartRevComp )
        .def_readonly( "sa_size", &Segment::iSaSize );

    py::class_<Seed>( m, "Seed" )
        .def_readwrite( "start", &Seed::uiStart )
        .def_readwrite( "size", &Seed::uiSize )
        .def_readwrite( "delta", &Seed::uiDelta )
        .def_readwrite( "start_ref", &Seed::uiPosOnReference )
        .def_readwrite( "on_forward_strand", &Seed::bOnForwStrand )
        .def_readwrite( "ambiguity", &Seed::uiAmbiguity );

    py::class_<Seeds>( m, "Seeds" )
        .def( "__len__", []( const Seeds& s ) { return s.vContent.size( ); } )
        .def( "__getitem__", []( const Seeds& s, size_t i ) { return s.vContent.at( i ); } )
        .def( "__iter__", []( const Seeds& s ) { return py::make_iterator( s.vContent.begin( ), s.vContent.end( ) ); },
              py::keep_alive<0, 1>( ) )
        .def_readonly( "index_of_strip", &Seeds::index_of_strip );

    py::enum_<MatchType>( m, "MatchType" )
        .value( "seed", MatchType::seed )
        .value( "match", MatchType::match )
        .value( "missmatch", MatchType::missmatch )
        .value( "insertion", MatchType::insertion )
        .value( "deletion", MatchType::deletion );

    py::class_<Alignment>( m, "Alignment" )
        .def( "begin_on_ref", []( const Alignment& a ) { return a.uiBeginOnRef; } )
        .def( "end_on_ref", []( const Alignment& a ) { return a.uiEndOnRef; } )
        .def( "__len__", []( const Alignment& a ) { return a.uiLength; } )
        .def( "length", []( const Alignment& a ) { return a.uiLength; } )
        .def( "get_score", &Alignment::score )
        .def( "num_seeds",
              []( const Alignment& a ) {
                  size_t n = 0;
                  for( auto& d : a.data )
                      n += d.first == MatchType::seed;
                  return n;
              } )
        .def_readonly( "data", &Alignment::data )
        .def_readonly( "begin_on_query", &Alignment::uiBeginOnQuery )
        .def_readonly( "end_on_query", &Alignment::uiEndOnQuery )
        .def_readonly( "index_of_strip", &Alignment::index_of_strip )
        .def_readonly( "mapping_quality", &Alignment::fMappingQuality )
        .def_readonly( "secondary", &Alignment::bSecondary )
        .def_readonly( "supplementary", &Alignment::bSupplementary )
        .def_readonly( "first", &Alignment::bFirst );

    py::class_<FMIndex>( m, "FMIndex" )
        .def( py::init<int>( ), py::arg( "device" ) = 0 )
        .def( "load", &FMIndex::vLoad ); // the reference's FMIndex(prefix) + Pack(prefix) files

    py::class_<BinarySeeding>( m, "BinarySeeding" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", &BinarySeeding::execute, py::arg( "fm_index" ), py::arg( "queries" ),
              py::arg( "max_segments" ) = 4096 )
        .def( "seed", &BinarySeeding::seed );
    py::class_<Harmonization>( m, "Harmonization" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", &Harmonization::execute );
    py::class_<NeedlemanWunsch>( m, "NeedlemanWunsch" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", []( NeedlemanWunsch& x, FMIndex& i, const std::vector<NucSeq>& q ) { return x.execute( i, q ); } );
    py::class_<MappingQuality>( m, "MappingQuality" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", []( MappingQuality& x, FMIndex& i, const std::vector<NucSeq>& q ) { return x.execute( i, q ); } );
    py::class_<PairedReads>( m, "PairedReads" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", []( PairedReads& x, FMIndex& i, const std::vector<NucSeq>& q ) { return x.execute( i, q ); } );
    // SmallInversions (export.cpp:58): alignments per read as MappingQuality returns them, the reads, the index
    py::class_<SmallInversions>( m, "SmallInversions" )
        .def( py::init<const ParameterSetManager&>( ), py::keep_alive<1, 2>( ) )
        .def( "execute", []( SmallInversions& x, FMIndex& i, const std::vector<std::vector<Alignment>>& a,
                             const std::vector<NucSeq>& q ) { return x.execute( i, a, q ); } );
    // FileReader (fileReader.h): all reads of a FASTA / FASTQ file
    m.def( "read_file", []( const std::string& sFile ) {
        ReadParser xParser( sFile );
        std::vector<NucSeq> v;
        NucSeq q;
        while( xParser.next( q ) )
            v.push_back( q );
        return v;
    } );
    // FileWriter / PairedFileWriter (fileWriter.h): the SAM text of a batch for the reads of a loaded index
    m.def( "sam_header", []( const FMIndex& i ) { return SamWriter( i.xContigs ).header( ); } );
    m.def( "sam_records", []( const FMIndex& i, const std::vector<NucSeq>& q, const std::vector<std::vector<Alignment>>& r,
                              bool bPaired ) {
        SamWriter xW( i.xContigs );
        std::string s;
        for( size_t u = 0; u < r.size( ); u++ )
            if( bPaired )
                xW.paired( s, q.at( 2 * u ), q.at( 2 * u + 1 ), r[ u ] );
            else
                xW.single( s, q.at( u ), r[ u ] );
        return s;
    }, py::arg( "fm_index" ), py::arg( "queries" ), py::arg( "records" ), py::arg( "paired" ) = false );
}
