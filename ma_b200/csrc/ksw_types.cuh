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

#define MA_KSW_RIGHT 0x02
#define MA_KSW_EXTZ_ONLY 0x40
#define MA_KSW_REV_CIGAR 0x80

} // namespace ma
