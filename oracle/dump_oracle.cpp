// CPU ORACLE — test infrastructure only (see oracle.h).
// Runs the restated path over a read file and writes the same MADUMP1 container as oracle/ref_dump.cpp, so that
// tests can compare oracle and compiled reference array by array.
#include "ma_oracle.h"
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <list>
#include <sstream>
#include <stdexcept>

using namespace oracle;

namespace
{
struct Dump
{
    std::list<std::pair<std::string, std::vector<int64_t>>> v;
    std::vector<int64_t>& arr( const std::string& n )
    {
        for( auto& p : v )
            if( p.first == n )
                return p.second;
        v.emplace_back( n, std::vector<int64_t>( ) );
        return v.back( ).second;
    }
    void write( const std::string& f )
    {
        std::ofstream o( f, std::ios::binary );
        o << "MADUMP1\n";
        for( auto& p : v )
        {
            o << p.first << " " << p.second.size( ) << "\n";
            o.write( (const char*)p.second.data( ), p.second.size( ) * 8 );
        }
    }
};
} // namespace

// stages: 1 = seeding, 2 = + extract, 3 = + SoC, 4 = + harmonization, 5 = + DP/alignments
// overrides: { bandwidth_ext, zdrop, padding, max_gap_area, min_bandwidth_gap }, -1 keeps the preset's value
static int g_aiOverride[ 5 ] = { -1, -1, -1, -1, -1 };
extern "C" void ma_oracle_set_overrides( const int* p )
{
    for( int i = 0; i < 5; i++ )
        g_aiOverride[ i ] = p ? p[ i ] : -1;
}
// parameters of the preset by the names of ma_oracle.h's Params ("name=value;name=value"), applied after the preset
static std::string g_sParamSet;
extern "C" void ma_oracle_set_params( const char* p )
{
    g_sParamSet = p ? p : "";
}
static void applyParamSet( Params& P )
{
    std::stringstream xS( g_sParamSet );
    std::string sItem;
    while( std::getline( xS, sItem, ';' ) )
    {
        const size_t uiEq = sItem.find( '=' );
        if( uiEq == std::string::npos )
            continue;
        const std::string n = sItem.substr( 0, uiEq );
        const double v = atof( sItem.c_str( ) + uiEq + 1 );
#define MA_SET( field )                                                                                                \
    if( n == #field )                                                                                                  \
    {                                                                                                                  \
        P.field = (decltype( P.field ))v;                                                                              \
        continue;                                                                                                      \
    }
        MA_SET( seeding_technique ) MA_SET( min_seed_length ) MA_SET( min_ambiguity ) MA_SET( max_ambiguity )
        MA_SET( seed_drop_min_size ) MA_SET( seed_drop_factor ) MA_SET( max_num_soc ) MA_SET( min_num_soc )
        MA_SET( soc_width ) MA_SET( soc_score_drop ) MA_SET( harm_score_min ) MA_SET( harm_score_min_rel )
        MA_SET( score_diff_tolerance ) MA_SET( max_score_lookahead ) MA_SET( switch_qlen ) MA_SET( max_delta_dist )
        MA_SET( min_delta_dist ) MA_SET( optimistic_gap_estimation ) MA_SET( gap_cost_cutting ) MA_SET( max_gap_area )
        MA_SET( genome_size_disable ) MA_SET( disable_heuristics ) MA_SET( padding ) MA_SET( bandwidth_ext )
        MA_SET( min_bandwidth_gap ) MA_SET( zdrop ) MA_SET( report_n ) MA_SET( min_alignment_score )
        MA_SET( max_supplementary_per_prim ) MA_SET( max_overlap_supplementary ) MA_SET( paired_mean )
        MA_SET( paired_std ) MA_SET( paired_bonus )
#undef MA_SET
        throw std::runtime_error( "ma_oracle_set_params: unknown parameter " + n );
    }
}
// "Minimum Genome Size for Heuristics" (parameter.h:873); negative keeps the preset's 10 M
static long long g_iMinGenomeSize = -1;
extern "C" void ma_oracle_set_min_genome_size( long long v )
{
    g_iMinGenomeSize = v;
}

extern "C" int ma_oracle_align_dump( const char* prefix, const char* reads_txt, const char* preset, const char* out,
                                     long long srand_base, int stages, char* err, int errcap )
{
    try
    {
        Index I;
        I.load( prefix );
        Params P;
        if( !P.preset( preset ) )
            throw std::runtime_error( "unknown preset" );
        if( g_iMinGenomeSize >= 0 )
            P.genome_size_disable = g_iMinGenomeSize;
        applyParamSet( P );
        if( g_aiOverride[ 0 ] >= 0 )
            P.bandwidth_ext = g_aiOverride[ 0 ];
        if( g_aiOverride[ 1 ] >= 0 )
            P.zdrop = g_aiOverride[ 1 ];
        if( g_aiOverride[ 2 ] >= 0 )
            P.padding = g_aiOverride[ 2 ];
        if( g_aiOverride[ 3 ] >= 0 )
            P.max_gap_area = g_aiOverride[ 3 ];
        if( g_aiOverride[ 4 ] >= 0 )
            P.min_bandwidth_gap = g_aiOverride[ 4 ];
        std::ifstream in( reads_txt );
        if( !in )
            throw std::runtime_error( "cannot open reads" );
        Dump D;
        auto& seg_off = D.arr( "seg_off" );
        auto& seg = D.arr( "seg" );
        auto& seed_off = D.arr( "seed_off" );
        auto& seed = D.arr( "seed" );
        auto& soc_off = D.arr( "soc_off" );
        auto& soc = D.arr( "soc" );
        auto& socseed_off = D.arr( "socseed_off" );
        auto& socseed = D.arr( "socseed" );
        auto& harm_off = D.arr( "harm_off" );
        auto& harmseed_off = D.arr( "harmseed_off" );
        auto& harmseed = D.arr( "harmseed" );
        auto& aln_off = D.arr( "aln_off" );
        auto& aln = D.arr( "aln" );
        auto& alndata_off = D.arr( "alndata_off" );
        auto& alndata = D.arr( "alndata" );
        auto& ksw_off = D.arr( "ksw_off" );
        auto& ksw_calls = D.arr( "ksw_calls" );
        auto& ksw_seq = D.arr( "ksw_seq" );
        auto& ksw_cigar = D.arr( "ksw_cigar" );
        auto& mq_off = D.arr( "mq_off" );
        auto& mq = D.arr( "mq" );
        auto& pr_off = D.arr( "pr_off" );
        auto& pr = D.arr( "pr" );
        mq_off.push_back( 0 ), pr_off.push_back( 0 );
        std::vector<Alignment> prevAlns;
        std::vector<MqAln> prevMq;
        int64_t prevQlen = 0;
        auto fBits = []( double d ) {
            int64_t i;
            memcpy( &i, &d, 8 );
            return i;
        };
        auto& work = D.arr( "work" ); // per read: n_ext
        for( auto* p : { &seg_off, &seed_off, &soc_off, &socseed_off, &harm_off, &harmseed_off, &aln_off,
                         &alndata_off, &ksw_off } )
            p->push_back( 0 );
        std::string line;
        size_t uiRead = 0;
        while( std::getline( in, line ) )
        {
            if( line.empty( ) )
                continue;
            std::vector<uint8_t> q;
            for( char c : line )
                q.push_back( c == 'A' || c == 'a' ? 0 : c == 'C' || c == 'c' ? 1 : c == 'G' || c == 'g' ? 2
                                                     : c == 'T' || c == 't' ? 3 : 4 );
            int64_t nExt = 0;
            auto segs = binary_seeding( I, P, q, &nExt );
            work.push_back( nExt );
            for( auto& s : segs )
            {
                int64_t a[ 5 ] = { s.start, s.size, s.sa.start, s.sa.rev, s.sa.size };
                seg.insert( seg.end( ), a, a + 5 );
            }
            seg_off.push_back( seg.size( ) / 5 );
            if( stages >= 2 )
            {
                auto seeds = extract_seeds( I, P, segs, (int64_t)q.size( ), nullptr );
                for( auto& s : seeds )
                {
                    int64_t a[ 6 ] = { s.q, s.len, s.r, s.amb, s.fw, s.delta };
                    seed.insert( seed.end( ), a, a + 6 );
                }
                seed_off.push_back( seed.size( ) / 6 );
                if( stages >= 3 )
                {
                    SoCQueue Q = strip_of_consideration( I, P, seeds, (int64_t)q.size( ) );
                    {
                        SoCQueue C = Q; // pop a copy completely to record the whole pop order
                        while( !C.maxima.empty( ) )
                        {
                            SoCOrder o = C.maxima.front( ).order;
                            unsigned idx;
                            auto pop = soc_pop( C, &idx );
                            int64_t a[ 4 ] = { (int64_t)o.acc_len, o.amb, o.count, idx };
                            soc.insert( soc.end( ), a, a + 4 );
                            for( auto& s : pop )
                            {
                                int64_t b[ 4 ] = { s.q, s.len, s.r, s.fw };
                                socseed.insert( socseed.end( ), b, b + 4 );
                            }
                            socseed_off.push_back( socseed.size( ) / 4 );
                        }
                    }
                    soc_off.push_back( soc.size( ) / 4 );
                    if( stages >= 4 )
                    {
                        if( srand_base >= 0 )
                            srand( (unsigned)( srand_base + uiRead ) );
                        auto sets = harmonization( I, P, Q, (int64_t)q.size( ) );
                        for( auto& S : sets )
                        {
                            for( auto& s : S.seeds )
                            {
                                int64_t b[ 5 ] = { s.q, s.len, s.r, s.fw, S.soc_index };
                                harmseed.insert( harmseed.end( ), b, b + 5 );
                            }
                            harmseed_off.push_back( harmseed.size( ) / 5 );
                        }
                        harm_off.push_back( harmseed_off.size( ) - 1 );
                        if( stages >= 5 )
                        {
                            std::vector<KswCall> log;
                            auto alns = needleman_wunsch( I, P, sets, q, &log );
                            for( auto& c : log )
                            {
                                ksw_calls.insert( ksw_calls.end( ), c.f, c.f + 16 );
                                ksw_seq.insert( ksw_seq.end( ), c.q.begin( ), c.q.end( ) );
                                ksw_seq.insert( ksw_seq.end( ), c.t.begin( ), c.t.end( ) );
                                ksw_cigar.insert( ksw_cigar.end( ), c.cigar.begin( ), c.cigar.end( ) );
                            }
                            ksw_off.push_back( ksw_calls.size( ) / 16 );
                            for( auto& A : alns )
                            {
                                int64_t a[ 8 ] = { A.begin_q, A.end_q,      A.begin_ref, A.end_ref,
                                                   A.score,   A.soc_index, A.length,    (int64_t)A.data.size( ) };
                                aln.insert( aln.end( ), a, a + 8 );
                                for( auto& d : A.data )
                                    alndata.push_back( d.first ), alndata.push_back( d.second );
                                alndata_off.push_back( alndata.size( ) / 2 );
                            }
                            aln_off.push_back( aln.size( ) / 8 );
                            // MappingQuality per read, PairedReads per pair of consecutive reads (as oracle/ref_dump.cpp)
                            auto vMq = mapping_quality( P, alns, (int64_t)q.size( ) );
                            for( auto& m : vMq )
                            {
                                int64_t a[ 3 ] = { m.idx, (int64_t)m.secondary | ( (int64_t)m.supplementary << 1 ),
                                                   fBits( m.mapq ) };
                                mq.insert( mq.end( ), a, a + 3 );
                            }
                            mq_off.push_back( mq.size( ) / 3 );
                            if( uiRead % 2 == 0 )
                                prevAlns = alns, prevMq = vMq, prevQlen = (int64_t)q.size( );
                            else
                            {
                                auto vPr = paired_reads( I, P, prevAlns, prevMq, prevQlen, alns, vMq, (int64_t)q.size( ) );
                                for( auto& x : vPr )
                                {
                                    int64_t a[ 4 ] = { x.mate, x.a.idx,
                                                       (int64_t)x.a.secondary | ( (int64_t)x.a.supplementary << 1 ),
                                                       fBits( x.a.mapq ) };
                                    pr.insert( pr.end( ), a, a + 4 );
                                }
                                pr_off.push_back( pr.size( ) / 4 );
                            }
                        }
                    }
                }
            }
            uiRead++;
        }
        D.write( out );
        return 0;
    }
    catch( const std::exception& e )
    {
        if( err && errcap > 0 )
            snprintf( err, errcap, "%s", e.what( ) );
        return -1;
    }
}
