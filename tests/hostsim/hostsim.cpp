// Host simulation of the device pipeline code (ma_b200/csrc/*.cuh compiled as plain C++) against the CPU oracle.
// TEST INFRASTRUCTURE: lets the host+device (MA_HD) routines be checked for bit-exactness without a GPU.
//   hostsim <index prefix> <reads.txt> <preset> [srand_base]
#include "../../ma_b200/csrc/nwglue.cuh"
#include "../../ma_b200/csrc/mapq.cuh"
#include "../../oracle/oracle.h"
#include "../../oracle/ma_oracle.h"
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <vector>

using namespace ma;

struct SetSink
{
    struct Set
    {
        std::vector<DSeed> seeds;
        unsigned soc;
    };
    std::vector<Set> v;
    void set( const DSeed* p, int n, unsigned soc )
    {
        v.push_back( Set{ std::vector<DSeed>( p, p + n ), soc } );
    }
    void pop_back( unsigned counter, unsigned minTries )
    {
        for( unsigned ui = 0; ui < counter && v.size( ) > minTries; ui++ )
            v.pop_back( );
    }
};

struct AllSegSink
{
    std::vector<SegRec> v;
    void seg( const SegRec& r )
    {
        v.push_back( r );
    }
};

struct MateState
{
    std::vector<DAln> al;
    std::vector<unsigned int> runs;
    std::vector<oracle::Alignment> oalns;
    std::vector<oracle::MqAln> omq;
    long long qlen = 0;
};

int main( int argc, char** argv )
{
    if( argc < 4 )
        return 2;
    oracle::Index OI;
    OI.load( argv[ 1 ] );
    oracle::Params OP;
    if( !OP.preset( argv[ 3 ] ) )
        return 2;
    // device-style index over host memory
    std::vector<long long> cs, cl;
    for( auto& c : OI.contigs )
        cs.push_back( c.start ), cl.push_back( c.length );
    std::vector<U4> bwt( ( OI.bwt.size( ) + 15 ) / 16 * 4 + 8 );
    { // reference blocks -> bit-plane blocks (what ma_b200_index_upload does on the device)
        std::vector<unsigned int> in( ( OI.bwt.size( ) + 15 ) / 16 * 16 + 16, 0 );
        memcpy( in.data( ), OI.bwt.data( ), OI.bwt.size( ) * 4 );
        for( size_t b = 0; b * 16 < OI.bwt.size( ); b++ )
            relayout_block( in.data( ) + 16 * b, (unsigned int*)bwt.data( ) + 16 * b );
    }
    DevIndex I;
    I.bwt = bwt.data( );
    I.sa = (const long long*)OI.sa.data( );
    I.pac = OI.pac.data( );
    I.contig_start = cs.data( ), I.contig_len = cl.data( );
    for( int i = 0; i < 5; i++ )
        I.L2[ i ] = (long long)OI.L2[ i ];
    I.primary = OI.primary, I.ref_len = OI.ref_len, I.fwd_len = OI.fwd_len;
    I.sa_intv = OI.sa_intv, I.n_contigs = (int)OI.contigs.size( );
    SeedParams SP{ OP.seeding_technique, OP.min_ambiguity, OP.max_ambiguity, OP.min_seed_length,
                   OP.seed_drop_min_size, OP.seed_drop_factor, OP.disable_heuristics, OP.genome_size_disable };

    HarmParams HP{ OP.match, OP.gap, OP.extend, OP.sv_penalty, OP.max_num_soc, OP.min_num_soc, OP.soc_width,
                   OP.rectangular_soc, OP.soc_score_drop, OP.harm_score_min, OP.harm_score_min_rel,
                   OP.score_diff_tolerance, OP.max_score_lookahead, OP.switch_qlen, OP.max_delta_dist,
                   OP.min_delta_dist, OP.optimistic_gap_estimation, OP.gap_cost_cutting, OP.disable_heuristics,
                   OP.genome_size_disable };
    const long long srandBase = argc > 4 ? atoll( argv[ 4 ] ) : 1;
    long nBadHarm = 0, nBadAln = 0;
    MapqParams MP{ OP.match, OP.report_n, OP.min_alignment_score, OP.max_supplementary_per_prim,
                   OP.max_overlap_supplementary, OP.paired_mean, OP.paired_std, OP.paired_bonus };
    NwParams NP{ OP.match, OP.mismatch, OP.gap, OP.extend, OP.sv_penalty, OP.max_gap_area, OP.padding,
                 OP.bandwidth_ext, OP.min_bandwidth_gap, OP.zdrop, 1 /* default scores: 1 x 1 gaps without DP */ };
    ma_oracle_score_t osc{ OP.match, OP.mismatch, OP.gap, OP.extend, OP.gap2, OP.extend2 };
    { // glibc rand() emulation check
        GlibcRand g;
        for( unsigned sd : { 0u, 1u, 77u, 123456789u, 4294967295u } )
        {
            srand( sd );
            g.seed( sd );
            for( int i = 0; i < 1000; i++ )
                if( rand( ) != g.next( ) )
                {
                    printf( "GlibcRand mismatch seed %u i %d\n", sd, i );
                    return 1;
                }
        }
    }
    std::ifstream in( argv[ 2 ] );
    std::string line;
    long nRead = 0, nBad = 0, nBadMq = 0;
    MateState prev;
    bool havePrev = false;
    unsigned long long nExtTotal = 0, nLookupTotal = 0;
    std::vector<SegRec> la( 600 ), lb( 600 );
    while( std::getline( in, line ) )
    {
        if( line.empty( ) )
            continue;
        std::vector<uint8_t> q;
        for( char c : line )
            q.push_back( c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4 );
        int64_t nExt = 0;
        auto oseg = oracle::binary_seeding( OI, OP, q, &nExt );
        AllSegSink sink;
        Seeder<AllSegSink> S( I, SP, q.data( ), (int)q.size( ), SeedScratch{ la.data( ), lb.data( ), 600 }, sink );
        S.run( );
        // drop-off heuristic applied by the caller of the seeder
        size_t sum = 0;
        for( auto& s : sink.v )
            sum += (size_t)s.size / (size_t)SP.drop_min_size;
        if( !SP.disable_heuristics && SP.drop_min_size != 0 && (double)sum < SP.drop_factor * (double)q.size( ) &&
            (unsigned long long)SP.genome_size_disable < (unsigned long long)I.ref_len )
            sink.v.clear( );
        bool ok = sink.v.size( ) == oseg.size( ) && !S.overflow;
        for( size_t i = 0; ok && i < oseg.size( ); i++ )
            ok = sink.v[ i ].start == oseg[ i ].start && sink.v[ i ].size == oseg[ i ].size &&
                 sink.v[ i ].sa.start == oseg[ i ].sa.start && sink.v[ i ].sa.rev == oseg[ i ].sa.rev &&
                 sink.v[ i ].sa.size == oseg[ i ].sa.size;
        if( S.nExt != nExt && oseg.size( ) )
            ok = false;
        { // the state-machine formulation used by seed_kernel must emit exactly the same segments
            AllSegSink sink2;
            int stk[ 80 ];
            // a short "shared" part (stride 3, like the interleaved device layout) so that both halves of SegList run
            U4 pk[ 3 * 6 ];
            int sz[ 3 * 6 ];
            unsigned short mu[ 3 * 6 ];
            SeederSM<AllSegSink> S2( I, SP, q.data( ), (int)q.size( ),
                                     SegList{ pk + 1, sz + 1, mu + 1, 3, 6, la.data( ), 606 }, sink2, stk );
            S2.run( );
            size_t sum2 = 0;
            for( auto& s : sink2.v )
                sum2 += (size_t)s.size / (size_t)SP.drop_min_size;
            if( !SP.disable_heuristics && SP.drop_min_size != 0 && (double)sum2 < SP.drop_factor * (double)q.size( ) &&
                (unsigned long long)SP.genome_size_disable < (unsigned long long)I.ref_len )
                sink2.v.clear( );
            nExtTotal += (unsigned long long)S2.nExt, nLookupTotal += (unsigned long long)S2.nLookup;
            bool ok2 = sink2.v.size( ) == sink.v.size( ) && !S2.overflow && ( S2.nExt == S.nExt );
            for( size_t i = 0; ok2 && i < sink.v.size( ); i++ )
                ok2 = sink2.v[ i ].start == sink.v[ i ].start && sink2.v[ i ].size == sink.v[ i ].size &&
                      sink2.v[ i ].sa.start == sink.v[ i ].sa.start && sink2.v[ i ].sa.rev == sink.v[ i ].sa.rev &&
                      sink2.v[ i ].sa.size == sink.v[ i ].sa.size;
            if( !ok2 )
                ok = false;
        }
        // locate
        auto oseeds = oracle::extract_seeds( OI, OP, oseg, (int64_t)q.size( ), nullptr );
        size_t k = 0;
        for( auto& s : sink.v )
        {
            if( (size_t)s.size < (size_t)SP.min_seed_len )
                continue;
            if( s.sa.size > SP.max_amb && SP.max_amb != 0 )
                continue;
            for( long long row = s.sa.start; row < s.sa.start + s.sa.size; row++, k++ )
            {
                long long r = bwt_sa( I, row, nullptr );
                bool fw = r < I.ref_len / 2;
                if( !fw )
                    r = I.ref_len - r - 1;
                long long delta = r + ( (long long)q.size( ) - s.start ) +
                                  ( (long long)q.size( ) + 1 ) * seq_id_for_position( I, r );
                if( k >= oseeds.size( ) || oseeds[ k ].r != r || oseeds[ k ].fw != fw || oseeds[ k ].delta != delta ||
                    oseeds[ k ].q != s.start || oseeds[ k ].len != s.size + 1 )
                    ok = false;
            }
        }
        if( k != oseeds.size( ) )
            ok = false;
        { // SoC + harmonization
            std::vector<DSeed> ds;
            for( auto& s : oseeds )
                ds.push_back( DSeed{ (int)s.q, (int)s.len, s.r, s.amb, s.fw ? 1 : 0, s.delta } );
            srand( (unsigned)( srandBase + nRead ) );
            auto Q = oracle::strip_of_consideration( OI, OP, oseeds, (int64_t)q.size( ) );
            auto osets = oracle::harmonization( OI, OP, Q, (int64_t)q.size( ) );
            SetSink ss;
            if( !ds.empty( ) )
            {
                std::vector<unsigned char> scratch( harm_scratch_need( ds.size( ) ) + 64 );
                HarmScratch W = harm_scratch_carve( scratch.data( ), ds.size( ) );
                soc_harm_read( I, HP, ds.data( ), (int)ds.size( ), (int)q.size( ), (unsigned)( srandBase + nRead ), W,
                               ss, 0 );
            }
            bool okh = ss.v.size( ) == osets.size( );
            for( size_t i = 0; okh && i < osets.size( ); i++ )
            {
                okh = ss.v[ i ].soc == osets[ i ].soc_index && ss.v[ i ].seeds.size( ) == osets[ i ].seeds.size( );
                for( size_t j = 0; okh && j < osets[ i ].seeds.size( ); j++ )
                {
                    auto& a = ss.v[ i ].seeds[ j ];
                    auto& b = osets[ i ].seeds[ j ];
                    okh = a.q == b.q && a.len == b.len && a.r == b.r && ( a.fw != 0 ) == b.fw;
                }
            }
            // ---- NW glue: plan -> DP (oracle ksw stands in for the kernel) -> assemble -> sort
            if( okh )
            {
                auto osetsCopy = osets;
                auto oalns = oracle::needleman_wunsch( OI, OP, osetsCopy, q, nullptr );
                struct Al
                {
                    DAln a;
                    std::vector<unsigned int> runs;
                };
                std::vector<Al> mine;
                for( auto& st : ss.v )
                {
                    Al al;
                    memset( &al.a, 0, sizeof( al.a ) );
                    al.a.soc_index = st.soc;
                    NwWindow w = nw_window( I, NP, st.seeds.data( ), (int)st.seeds.size( ) );
                    if( w.valid )
                    {
                        NwPlanner cnt( NP, nullptr, 0, (long long)w.beginRef );
                        nw_walk( st.seeds.data( ), (int)st.seeds.size( ), (int)q.size( ), w, cnt );
                        std::vector<KswTask> tasks( cnt.n + 1 );
                        NwPlanner pl( NP, tasks.data( ), 0, (long long)w.beginRef );
                        nw_walk( st.seeds.data( ), (int)st.seeds.size( ), (int)q.size( ), w, pl );
                        std::vector<KswOut> res( pl.n + 1 );
                        std::vector<unsigned int> cig;
                        for( int t = 0; t < pl.n; t++ )
                        {
                            const KswTask& T = tasks[ t ];
                            std::vector<uint8_t> tq( T.qlen ), tt( T.tlen );
                            for( int i = 0; i < T.qlen; i++ )
                                tq[ i ] = q[ T.qoff + ( ( T.tag & MA_TASK_QREV ) ? -i : i ) ];
                            for( int i = 0; i < T.tlen; i++ )
                                tt[ i ] = (uint8_t)pack_virtual( I, T.toff + ( ( T.tag & MA_TASK_TREV ) ? -i : i ) );
                            ma_oracle_ksw_t ez;
                            std::vector<uint32_t> c( T.qlen + T.tlen + 8 );
                            int64_t cells;
                            ma_oracle_ksw( T.qlen, tq.data( ), T.tlen, tt.data( ), &osc, T.w, T.zdrop, T.flag, &ez,
                                           c.data( ), (int)c.size( ), &cells );
                            KswOut& o = res[ t ];
                            o.max = ez.max, o.zdropped = ez.zdropped, o.max_q = ez.max_q, o.max_t = ez.max_t;
                            o.mqe = ez.mqe, o.mqe_t = ez.mqe_t, o.mte = ez.mte, o.mte_q = ez.mte_q, o.score = ez.score;
                            o.n_cigar = ez.n_cigar, o.reach_end = ez.reach_end, o.cigar_off = (long long)cig.size( );
                            cig.insert( cig.end( ), c.begin( ), c.begin( ) + ez.n_cigar );
                        }
                        std::vector<unsigned int> runs( 2 * q.size( ) + 4096 );
                        NwAssembler as( I, NP, q.data( ), w.beginRef, res.data( ), cig.data( ), runs.data( ),
                                        (int)runs.size( ) );
                        nw_walk( st.seeds.data( ), (int)st.seeds.size( ), (int)q.size( ), w, as );
                        as.removeDangeling( );
                        if( as.overflow || as.next != pl.n )
                            okh = false;
                        al.a.begin_ref = (long long)as.beginR, al.a.end_ref = (long long)as.endR;
                        al.a.begin_q = (int)as.beginQ, al.a.end_q = (int)as.endQ, al.a.score = as.score;
                        al.a.length = (int)as.length;
                        al.runs.assign( runs.begin( ) + as.front, runs.begin( ) + as.nRuns );
                    }
                    mine.push_back( al );
                }
                // final std::sort with Alignment::larger (needlemanWunsch.h:131-132)
                std::vector<int> ord( mine.size( ) );
                for( size_t i = 0; i < ord.size( ); i++ )
                    ord[ i ] = (int)i;
                stl::sort( ord.data( ), ord.data( ) + ord.size( ), [ & ]( int a, int b ) {
                    if( mine[ a ].a.score == mine[ b ].a.score )
                        return mine[ a ].a.soc_index < mine[ b ].a.soc_index;
                    return mine[ a ].a.score > mine[ b ].a.score;
                } );
                bool oka = okh && mine.size( ) == oalns.size( );
                for( size_t i = 0; oka && i < oalns.size( ); i++ )
                {
                    const Al& m = mine[ ord[ i ] ];
                    const auto& o = oalns[ i ];
                    oka = m.a.begin_q == o.begin_q && m.a.end_q == o.end_q && m.a.begin_ref == o.begin_ref &&
                          m.a.end_ref == o.end_ref && m.a.score == o.score && m.a.soc_index == o.soc_index &&
                          m.a.length == o.length && m.runs.size( ) == o.data.size( );
                    for( size_t j = 0; oka && j < o.data.size( ); j++ )
                        oka = (int)( m.runs[ j ] & 7 ) == o.data[ j ].first && (long long)( m.runs[ j ] >> 3 ) == o.data[ j ].second;
                }
                if( !oka )
                {
                    if( nBadAln < 5 )
                        printf( "read %ld ALN MISMATCH (%zu vs %zu)\n", nRead, mine.size( ), oalns.size( ) );
                    nBadAln++;
                }
                // ---- MappingQuality per read, PairedReads per pair of consecutive reads (mapq.cuh vs oracle)
                if( oka )
                {
                    MateState cur;
                    cur.qlen = (long long)q.size( );
                    cur.oalns = oalns;
                    for( size_t i = 0; i < mine.size( ); i++ )
                    { // slab order = set order; rank = position in the NeedlemanWunsch result
                        DAln a = mine[ i ].a;
                        a.run_off = (long long)cur.runs.size( ), a.n_runs = (int)mine[ i ].runs.size( );
                        cur.runs.insert( cur.runs.end( ), mine[ i ].runs.begin( ), mine[ i ].runs.end( ) );
                        cur.al.push_back( a );
                    }
                    for( size_t i = 0; i < ord.size( ); i++ )
                        cur.al[ ord[ i ] ].rank = (int)i;
                    std::vector<int> sc( cur.al.size( ) + 1 );
                    const int nRep = mapping_quality_read( MP, cur.al.data( ), (int)cur.al.size( ), cur.runs.data( ),
                                                           cur.qlen, sc.data( ) );
                    cur.omq = oracle::mapping_quality( OP, oalns, cur.qlen );
                    bool okm = nRep == (int)cur.omq.size( );
                    for( size_t k = 0; okm && k < cur.omq.size( ); k++ )
                    {
                        okm = false;
                        for( auto& a : cur.al )
                            if( a.rank_mq == (int)k )
                                okm = a.rank == cur.omq[ k ].idx &&
                                      ( ( a.flags & 1 ) != 0 ) == cur.omq[ k ].secondary &&
                                      ( ( a.flags & 2 ) != 0 ) == cur.omq[ k ].supplementary &&
                                      memcmp( &a.mapq, &cur.omq[ k ].mapq, 8 ) == 0;
                    }
                    if( okm && ( nRead & 1 ) && havePrev )
                    {
                        // the two mates' records live in one slab on the device: concatenate
                        std::vector<DAln> both = prev.al;
                        std::vector<unsigned int> runs = prev.runs;
                        for( DAln a : cur.al )
                            a.run_off += (long long)prev.runs.size( ), both.push_back( a );
                        runs.insert( runs.end( ), cur.runs.begin( ), cur.runs.end( ) );
                        const int n1 = (int)prev.al.size( ), n2 = (int)cur.al.size( );
                        const int cap = std::max( 1, n1 * n2 );
                        std::vector<int> o1( n1 + 1 ), o2( n2 + 1 ), meta( 2 * cap );
                        std::vector<long long> scs( cap );
                        const int nOut = paired_reads_pair( MP, I.ref_len, both.data( ), n1, prev.qlen, both.data( ) + n1,
                                                            n2, cur.qlen, runs.data( ), o1.data( ), o2.data( ),
                                                            scs.data( ), meta.data( ), cap );
                        auto omq1 = prev.omq, omq2 = cur.omq;
                        auto opr = oracle::paired_reads( OI, OP, prev.oalns, omq1, prev.qlen, oalns, omq2, cur.qlen );
                        okm = nOut == (int)opr.size( );
                        for( size_t k = 0; okm && k < opr.size( ); k++ )
                        {
                            okm = false;
                            for( int i = 0; i < n1 + n2; i++ )
                                if( both[ i ].pair_rank == (int)k && ( i >= n1 ) == ( opr[ k ].mate == 1 ) )
                                    okm = both[ i ].rank == opr[ k ].a.idx &&
                                          ( ( both[ i ].flags & 1 ) != 0 ) == opr[ k ].a.secondary &&
                                          ( ( both[ i ].flags & 2 ) != 0 ) == opr[ k ].a.supplementary &&
                                          memcmp( &both[ i ].mapq, &opr[ k ].a.mapq, 8 ) == 0;
                        }
                    }
                    if( !okm )
                    {
                        if( nBadMq < 5 )
                            printf( "read %ld MAPQ/PAIR MISMATCH\n", nRead );
                        nBadMq++;
                    }
                    prev = cur, havePrev = true;
                }
                else
                    havePrev = false;
            }
            if( !okh )
            {
                if( nBadHarm < 5 )
                    printf( "read %ld HARM MISMATCH (sets %zu vs %zu)\n", nRead, ss.v.size( ), osets.size( ) );
                nBadHarm++;
            }
        }
        if( !ok )
        {
            if( nBad < 5 )
                printf( "read %ld MISMATCH (segs %zu vs %zu)\n", nRead, sink.v.size( ), oseg.size( ) );
            nBad++;
        }
        nRead++;
    }
    printf( "hostsim: %ld reads, seeding mismatches %ld, soc/harm mismatches %ld, alignment mismatches %ld\n", nRead,
            nBad, nBadHarm, nBadAln );
    printf( "hostsim: mapping quality / pairing mismatches %ld\n", nBadMq );
    printf( "hostsim: %llu extensions, %llu occurrence-table lookups (memo hit rate %.1f%%)\n", nExtTotal, nLookupTotal,
            nExtTotal ? 100.0 * (double)( nExtTotal - nLookupTotal ) / (double)nExtTotal : 0.0 );
    return ( nBad || nBadHarm || nBadAln || nBadMq ) ? 1 : 0;
}
