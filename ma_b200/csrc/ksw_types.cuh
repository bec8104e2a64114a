// Plain data types of the banded-DP batch interface, shared by the kernel (ksw.cuh), the NW glue (nwglue.cuh)
// and the host simulation. Mirrors kswcpp_extz_t / the kswcpp_dispatch arguments (libs/kswcpp/inc/kswcpp.h:31-41,165-190).
#pragma once
#include "stl_exact.cuh"

namespace ma
{

struct KswTask
{
    long long qoff, toff; // byte offsets into the sequence slab
    int qlen, tlen, w, zdrop, flag, tag;
};

struct KswOut
{
    int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end, status;
    long long cigar_off; // word offset into the cigar slab
    long long cells; // band cells processed (st0..en0 over all rows) — the GCUPS unit
};

struct KswScore
{
    int match, mismatch; // mismatch as a (negative) score
    int q, e, q2, e2; // after the q/q2 swap of kswcpp_core.h:367-375
    int qe_row0; // q + e BEFORE that swap: the reference's scalar qe (kswcpp_core.h:338) enters H[0] of row 0 (:247)
    int long_thres, long_diff;
    int min16; // iOverallMinScr (negative) for the int16/int32 switch, kswcpp.h:101-115
    int early_return; // -min_sc > 2(q+e): the reference returns right after ksw_reset_extz
};

// task addressing modes (KswTask::tag)
#define MA_TASK_QREV 1 /* query element i is at qoff - i */
#define MA_TASK_TREV 2 /* target element i is at toff - i */
#define MA_TASK_TPACK 4 /* target lives in the pack: toff is a position in the virtual forward+reverse text */
#define MA_TASK_EARLYSTOP 8 /* only max / max_q / max_t / CIGAR are consumed: stop once they are provably final */
#define MA_TASK_SKIP 16 /* pipeline: the task's seed set was given up (band beyond the largest window), do not run it */

#define MA_KSW_RIGHT 0x02
#define MA_KSW_EXTZ_ONLY 0x40
#define MA_KSW_REV_CIGAR 0x80

// Where the two sequences of a problem live. Standalone batches (ma_b200_ksw_batch) read both from a byte slab;
// the alignment pipeline reads the query from the read slab (forwards or backwards) and the target straight from
// the 2-bit pack through its virtual forward+reverse-complement text, so no reference window is ever materialised.
struct SeqAccess
{
    const unsigned char* qbase;
    long long qoff;
    int qstep;
    const unsigned char* tslab;
    const unsigned char* pac;
    long long fwd_len;
    long long toff;
    int tstep;
    MA_HD inline int Q( long long i ) const
    {
        return qbase[ qoff + qstep * i ];
    }
    MA_HD inline int T( long long i ) const
    {
        const long long p = toff + tstep * i;
        if( pac == nullptr )
            return tslab[ p ];
        const long long f = p < fwd_len ? p : 2 * fwd_len - 1 - p;
        const int b = pac[ f >> 2 ] >> ( ( ~f & 3 ) << 1 ) & 3;
        return p < fwd_len ? b : 3 - b;
    }
};

// scoring parameters for which the reference's int8 difference arithmetic can never wrap (see ksw.cuh, packed path)
MA_HD inline bool ksw_p2_params_ok( const KswScore& P )
{
    const int Q = P.q + P.e > P.q2 + P.e2 ? P.q + P.e : P.q2 + P.e2;
    const int mis = -P.mismatch > P.e2 ? -P.mismatch : P.e2;
    const int gq = P.q > P.q2 ? P.q : P.q2;
    const int ld = P.long_diff < 0 ? -P.long_diff : P.long_diff;
    return P.match > 0 && P.q >= 0 && P.e >= 0 && P.q2 >= 0 && P.e2 >= 0 && P.mismatch <= 0 &&
           2 * Q + P.match + mis + gq + ld <= 127;
}


// aligned band width in cells: n_col_ * 16 (kswcpp_core.h:401-402)
MA_HD inline int ksw_ncol16( int qlen, int tlen, int w )
{
    if( w < 0 )
        w = tlen > qlen ? tlen : qlen;
    int n = qlen < tlen ? qlen : tlen;
    n = n < w + 1 ? n : w + 1;
    return ( ( n + 15 ) / 16 + 1 ) * 16;
}


// lane 0 only. Walks the traceback slab (kswcpp_core.h:76-150, is_rot = 1, min_intron_len = 0) and pushes run-length
// ops in backtrack order into cig[]; returns the number of words or -1 on overflow.
MA_HD inline int ksw_backtrack( const unsigned char* tb, int ncol16, int qlen, int tlen, int w, int i0, int j0,
                                     unsigned int* cig, int cap )
{
    int i = i0, j = j0; // qlen + tlen < 2^31
    int state = 0, n = 0;
    unsigned int cur = 0; // current run: len<<4|op, 0 = none
    auto push = [ & ]( unsigned int op, unsigned int len ) {
        if( cur != 0 && ( cur & 0xf ) == op )
            cur += len << 4;
        else
        {
            if( cur != 0 )
            {
                if( n < cap )
                    cig[ n ] = cur;
                n++;
            }
            cur = len << 4 | op;
        }
    };
    while( i >= 0 && j >= 0 )
    {
        const int r = i + j;
        // band limits of row r (kswcpp_core.h:541-548)
        int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0;
        if( st0 < ( ( r - w + 1 ) >> 1 ) )
            st0 = ( r - w + 1 ) >> 1;
        int en0 = tlen - 1 < r ? tlen - 1 : r;
        if( en0 > ( ( r + w ) >> 1 ) )
            en0 = ( r + w ) >> 1;
        const int off = st0 & ~15, off_end = en0 | 15;
        int force_state = -1;
        if( i < off )
            force_state = 2;
        if( i > off_end )
            force_state = 1;
        unsigned int tmp = force_state < 0 ? tb[ (long long)r * ncol16 + ( i - off ) ] : 0;
        if( state == 0 )
            state = tmp & 7;
        else if( !( tmp >> ( state + 2 ) & 1 ) )
            state = 0;
        if( state == 0 )
            state = tmp & 7;
        if( force_state >= 0 )
            state = force_state;
        if( state == 0 )
            push( 0, 1 ), --i, --j;
        else if( state == 1 || state == 3 )
            push( 2, 1 ), --i;
        else
            push( 1, 1 ), --j;
    }
    if( i >= 0 )
        push( 2, (unsigned int)i + 1 );
    if( j >= 0 )
        push( 1, (unsigned int)j + 1 );
    if( cur != 0 )
    {
        if( n < cap )
            cig[ n ] = cur;
        n++;
    }
    return n > cap ? -1 : n;
}


// decide where the backtrack starts (kswcpp_core.h:796-835); returns false if there is no backtrack
MA_HD inline bool ksw_bt_start( KswOut& ez, int qlen, int tlen, int flag, int& i0, int& j0 )
{
    if( !ez.zdropped && !( flag & MA_KSW_EXTZ_ONLY ) )
    {
        i0 = tlen - 1, j0 = qlen - 1;
        return true;
    }
    if( !ez.zdropped && ( flag & MA_KSW_EXTZ_ONLY ) && ez.mqe > ez.max )
    {
        ez.reach_end = 1;
        i0 = ez.mqe_t, j0 = qlen - 1;
        return true;
    }
    if( ez.max_t >= 0 && ez.max_q >= 0 )
    {
        i0 = ez.max_t, j0 = ez.max_q;
        return true;
    }
    return false;
}


} // namespace ma
