// CPU ORACLE — test infrastructure only (see oracle.h).  Internal C++ types of the restatement.
#pragma once
#include "oracle.h"
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

namespace oracle
{

// ParameterSetManager preset + pGlobalParams values the path reads (libs/ms/inc/ms/util/parameter.h)
struct Params
{
    int match = 2, mismatch = 4, gap = 4, extend = 2, gap2 = 24, extend2 = 1, sv_penalty = 100;
    int seeding_technique = 0, min_seed_length = 16, min_ambiguity = 0, max_ambiguity = 100;
    int seed_drop_min_size = 15;
    double seed_drop_factor = 0.005;
    int max_num_soc = 30, min_num_soc = 1, soc_width = 0, rectangular_soc = 1;
    double soc_score_drop = 0.1;
    int harm_score_min = 18;
    double harm_score_min_rel = 0.002, score_diff_tolerance = 0.0001;
    int max_score_lookahead = 3, switch_qlen = 800;
    double max_delta_dist = 0.1;
    int min_delta_dist = 16;
    int optimistic_gap_estimation = 1, gap_cost_cutting = 1, max_gap_area = 20;
    int64_t genome_size_disable = 10000000;
    int disable_heuristics = 0;
    int padding = 1000, bandwidth_ext = 512, min_bandwidth_gap = 20, zdrop = 200;
    // MappingQuality / PairedReads (parameter.h:652-668, 721-737)
    int report_n = 0, min_alignment_score = 75, max_supplementary_per_prim = 1;
    double max_overlap_supplementary = 0.1, paired_mean = 400, paired_std = 150, paired_bonus = 1.25;
    int use_paired_reads = 0;
    bool preset( std::string sName );
};

struct Contig
{
    std::string name;
    int64_t start, length;
};

// FMIndex + Pack as loaded from the reference's on-disk files
struct Index
{
    int64_t primary = 0;
    uint64_t L2[ 6 ] = { 0, 0, 0, 0, 0, 0 };
    std::vector<uint32_t> bwt; // 64-byte blocks: 4 x u64 counts + 8 x u32 of 16 symbols
    std::vector<int64_t> sa; // one sample per sa_intv rows, sa[0] = -1
    int sa_intv = 32;
    int64_t ref_len = 0; // forward + reverse
    std::vector<uint8_t> pac; // 2 bit per base, forward strand
    int64_t fwd_len = 0;
    std::vector<Contig> contigs;
    void load( const std::string& sPrefix );

    void occ4( int64_t k, int64_t cnt[ 4 ] ) const;
    int64_t occ( int64_t k, int c ) const;
    int B0( int64_t k ) const;
    int64_t invPsi( int64_t k ) const;
    int64_t bwt_sa( int64_t k ) const;
    // pack
    int nuc( int64_t pos ) const
    {
        return pac[ pos >> 2 ] >> ( ( ~pos & 3 ) << 1 ) & 3;
    }
    bool onReverse( int64_t p ) const
    {
        return p >= fwd_len;
    }
    int64_t seqIdForPosition( int64_t pos ) const;
    int64_t seqIdForPositionOrRev( int64_t pos ) const;
    int64_t startOfSeqOrRev( int64_t id ) const;
    int64_t endOfSeqOrRev( int64_t id ) const;
    bool bridging( int64_t begin, int64_t size ) const;
    void extract( int64_t begin, int64_t end, std::vector<uint8_t>& out ) const;
};

struct SAInterval
{
    int64_t start = 0, rev = 0, size = 0;
    SAInterval revComp( ) const
    {
        return SAInterval{ rev, start, size };
    }
    int64_t end( ) const
    {
        return start + size;
    }
};

struct Segment
{
    int64_t start, size; // size = length - 1
    SAInterval sa;
    int64_t end( ) const
    {
        return start + size;
    }
};

struct Seed
{
    int64_t q = 0, len = 0, r = 0;
    unsigned amb = 0;
    bool fw = true;
    int64_t delta = 0;
    uint64_t soc_nt = 0;
    int64_t end( ) const
    {
        return q + len;
    }
    int64_t end_ref( ) const
    {
        return r + len;
    }
};

struct SoCOrder
{
    uint64_t acc_len = 0;
    unsigned amb = 0, count = 0;
    void add( const Seed& s )
    {
        amb += s.amb, count++, acc_len += s.len;
    }
    void sub( const Seed& s )
    {
        amb -= s.amb, acc_len -= s.len, count--;
    }
    bool operator<( const SoCOrder& o ) const
    {
        if( acc_len == o.acc_len )
            return amb > o.amb;
        return acc_len < o.acc_len;
    }
};

struct SoC
{
    SoCOrder order;
    size_t begin, end; // indices into the seed vector
};

struct SoCQueue
{
    std::vector<Seed> seeds;
    std::vector<SoC> maxima;
    unsigned next_index = 0;
};

struct SeedSet
{
    std::vector<Seed> seeds;
    unsigned soc_index = 0;
};

enum MatchType
{
    MT_SEED = 0,
    MT_MATCH = 1,
    MT_MISSMATCH = 2,
    MT_INSERTION = 3,
    MT_DELETION = 4
};

struct Alignment
{
    std::vector<std::pair<int, int64_t>> data;
    int64_t length = 0, begin_ref = 0, end_ref = 0, begin_q = 0, end_q = 0, score = 0;
    unsigned soc_index = 0;
};

struct KswCall
{
    int64_t f[ 16 ];
    std::vector<uint8_t> q, t;
    std::vector<uint32_t> cigar;
};

SAInterval extend_backward( const Index& I, const SAInterval& ik, int c );
SAInterval init_interval( const Index& I, int c );
std::vector<Segment> binary_seeding( const Index& I, const Params& P, const std::vector<uint8_t>& q, int64_t* pnExt );
std::vector<Seed> extract_seeds( const Index& I, const Params& P, const std::vector<Segment>& segs, int64_t qlen,
                                 int64_t* pnInvPsi );
SoCQueue strip_of_consideration( const Index& I, const Params& P, std::vector<Seed> seeds, int64_t qlen );
std::vector<Seed> soc_pop( SoCQueue& Q, unsigned* pIndex );
std::vector<SeedSet> harmonization( const Index& I, const Params& P, SoCQueue& Q, int64_t qlen );
std::vector<Alignment> needleman_wunsch( const Index& I, const Params& P, std::vector<SeedSet>& sets,
                                         std::vector<uint8_t>& query, std::vector<KswCall>* pLog );

// one element of MappingQuality's result: an alignment of the NeedlemanWunsch result with its flags and quality
struct MqAln
{
    int idx = 0;
    bool secondary = false, supplementary = false;
    double mapq = NAN;
};
struct PairAln
{
    int mate; // 0 / 1
    MqAln a;
};
std::vector<MqAln> mapping_quality( const Params& P, const std::vector<Alignment>& alns, int64_t qlen );
std::vector<PairAln> paired_reads( const Index& I, const Params& P, const std::vector<Alignment>& alns1,
                                   std::vector<MqAln>& mq1, int64_t qlen1, const std::vector<Alignment>& alns2,
                                   std::vector<MqAln>& mq2, int64_t qlen2 );

} // namespace oracle
