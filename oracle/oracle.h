/* CPU ORACLE — test infrastructure only.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this library.
 * The product (ma_b200/csrc, libma_b200.so) never includes this header and never links liboracle.
 *
 * Plain C++ restatement of the reference's read-alignment hot path (ITBE-Lab/ma); each function cites the
 * reference file:line it follows.  Pinned against the compiled reference (oracle/_ref, built by oracle/Makefile
 * from the unmodified sources) and against the golden dumps under tests/golden/.
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* kswcpp flags, libs/kswcpp/inc/kswcpp.h:21-28 */
#define MA_KSW_SCORE_ONLY 0x01
#define MA_KSW_RIGHT 0x02
#define MA_KSW_EXTZ_ONLY 0x40
#define MA_KSW_REV_CIGAR 0x80

/* pGlobalParams scoring, libs/ms/inc/ms/util/parameter.h:1032-1046 (penalties are positive numbers) */
typedef struct
{
    int match, mismatch, gap, extend, gap2, extend2;
} ma_oracle_score_t;

/* kswcpp_extz_t without the cigar pointer, libs/kswcpp/inc/kswcpp.h:31-41 */
typedef struct
{
    int max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end;
} ma_oracle_ksw_t;

/* kswcpp_dispatch (kswcpp.h:165-190). cigar words = len<<4 | op (0 M, 1 I, 2 D). cells = band cells processed. */
int ma_oracle_ksw( int qlen, const uint8_t* query, int tlen, const uint8_t* target, const ma_oracle_score_t* sc, int w,
                   int zdrop, int flag, ma_oracle_ksw_t* ez, uint32_t* cigar, int cigar_cap, int64_t* cells );

#ifdef __cplusplus
}
#endif
