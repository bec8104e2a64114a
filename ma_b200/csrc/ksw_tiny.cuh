// Gap fills between seeds (NeedlemanWunsch::dynPrg -> ksw(), needlemanWunsch.cpp:111-168, 556-573): global alignments
// of at most "Maximal Gap Size" (20) bases on either side whose band (w = max(20, |tlen - qlen| + 10)) covers the whole
// matrix, and of which the caller reads nothing but the CIGAR. 550 k such problems of ~100 cells per 2 M Illumina reads:
// a warp per problem (ksw_batch_kernel) spends its time on set-up and row bookkeeping (13 GCUPS), so here ONE THREAD
// runs one problem: the difference recurrence of kswcpp_inner_core (kswcpp_core.h:640-760, left-aligned gaps) over the
// cells of the matrix only, state in thread-local arrays, the walk of ksw_backtrack (kswcpp_core.h:76-150) from the
// bottom-right corner.
//
// Why only the matrix cells: with w >= max(qlen, tlen) the band limits of every anti-diagonal are the matrix borders
// (t - i <= tlen - 1 <= w - 1 and i - t <= qlen - 1 <= w - 1). The reference additionally computes the cells that pad its
// 16-aligned column ranges; as in the FAST mode of ksw_rows (ksw.cuh) they feed no cell of the matrix: a column's
// u / y / y2 are re-initialised when it enters (kswcpp_core.h:562-579), and x / v / x2 are read from the left
// neighbour's value of the PREVIOUS anti-diagonal, which is a matrix cell.
#pragma once
#include "ksw_types.cuh"

namespace ma
{

#define MA_TINY_MAX 32

// does ksw_tiny_kernel take this task? (pipeline tasks only: the standalone batch interface promises every field)
MA_HD inline bool ksw_tiny_ok( const KswScore& P, int qlen, int tlen, int w, int flag, int tag )
{
    if( flag != 0 || ( tag & ( MA_TASK_EARLYSTOP | MA_TASK_QREV | MA_TASK_TREV ) ) || P.early_return )
        return false;
    if( qlen < 1 || tlen < 1 || qlen > MA_TINY_MAX || tlen > MA_TINY_MAX )
        return false;
    return w < 0 || w >= ( qlen > tlen ? qlen : tlen );
}

__device__ __forceinline__ int tiny_w8( int x )
{
    return (int)(signed char)x;
}

__global__ void __launch_bounds__( 128 ) ksw_tiny_kernel( KswBatchArgs A )
{
    const KswScore P = A.score;
    const int q = P.q, e = P.e, q2 = P.q2, e2 = P.e2, qe = q + e, qe2 = q2 + e2;
    const int init6 = tiny_w8( -q - e ), init25 = tiny_w8( -q2 - e2 );
    const int scN = -e2, scM = P.match, scX = P.mismatch;
    unsigned long long cellsLocal = 0;
    for( int slot = blockIdx.x * blockDim.x + threadIdx.x; slot < A.n; slot += gridDim.x * blockDim.x )
    {
        const int ti = A.order[ slot ];
        const KswTask T = A.tasks[ ti ];
        SeqAccess sa;
        sa.qbase = A.seq, sa.qoff = T.qoff, sa.qstep = 1;
        sa.tslab = A.seq, sa.toff = T.toff, sa.tstep = 1;
        sa.pac = ( T.tag & MA_TASK_TPACK ) ? A.pac : nullptr, sa.fwd_len = A.fwd_len;
        const int qlen = T.qlen, tlen = T.tlen;
        signed char u[ MA_TINY_MAX ], v[ MA_TINY_MAX ], x[ MA_TINY_MAX ], y[ MA_TINY_MAX ], x2[ MA_TINY_MAX ], y2[ MA_TINY_MAX ];
        unsigned char tc[ MA_TINY_MAX ], qc[ MA_TINY_MAX ];
        unsigned char tbm[ MA_TINY_MAX * MA_TINY_MAX ]; // [query row][target column]
        for( int t = 0; t < tlen; t++ )
        {
            u[ t ] = v[ t ] = x[ t ] = y[ t ] = (signed char)init6;
            x2[ t ] = y2[ t ] = (signed char)init25;
            tc[ t ] = (unsigned char)sa.T( t );
        }
        for( int i = 0; i < qlen; i++ )
            qc[ i ] = (unsigned char)sa.Q( i );
        const int nrows = qlen + tlen - 1;
        for( int r = 0; r < nrows; r++ )
        {
            const int st0 = r - qlen + 1 > 0 ? r - qlen + 1 : 0, en0 = r < tlen - 1 ? r : tlen - 1;
            const int first_col = tiny_w8( r == 0 ? -q - e : r < P.long_thres ? -e : r == P.long_thres ? P.long_diff : -e2 );
            if( en0 == r )
                y[ r ] = (signed char)init6, y2[ r ] = (signed char)init25, u[ r ] = (signed char)first_col;
            // in place, from the last column down: a cell reads x / v / x2 of column t - 1 as the previous row left them
            for( int t = en0; t >= st0; t-- )
            {
                const int xt1 = t > 0 ? x[ t - 1 ] : init6, vt1 = t > 0 ? v[ t - 1 ] : first_col;
                const int x2t1 = t > 0 ? x2[ t - 1 ] : init25;
                const int ut = u[ t ], yo = y[ t ], y2o = y2[ t ];
                const int ca = tc[ t ], cb = qc[ r - t ];
                int z = ( ca == 4 || cb == 4 ) ? scN : ( ca == cb ? scM : scX );
                int a = tiny_w8( xt1 + vt1 ), b = tiny_w8( yo + ut ), a2 = tiny_w8( x2t1 + vt1 ), b2 = tiny_w8( y2o + ut );
                int d = a > z ? 1 : 0; // left-aligned gaps (flag 0), kswcpp_core.h:668-691
                z = max( z, a );
                d = b > z ? 2 : d;
                z = max( z, b );
                d = a2 > z ? 3 : d;
                z = max( z, a2 );
                d = b2 > z ? 4 : d;
                z = max( z, b2 );
                z = min( z, scM );
                const int un = tiny_w8( z - vt1 ), vn = tiny_w8( z - ut );
                int tmp = tiny_w8( z - q );
                a = tiny_w8( a - tmp ), b = tiny_w8( b - tmp );
                tmp = tiny_w8( z - q2 );
                a2 = tiny_w8( a2 - tmp ), b2 = tiny_w8( b2 - tmp );
                d |= a > 0 ? 0x08 : 0;
                d |= b > 0 ? 0x10 : 0;
                d |= a2 > 0 ? 0x20 : 0;
                d |= b2 > 0 ? 0x40 : 0;
                u[ t ] = (signed char)un, v[ t ] = (signed char)vn;
                x[ t ] = (signed char)( max( a, 0 ) - qe ), y[ t ] = (signed char)( max( b, 0 ) - qe );
                x2[ t ] = (signed char)( max( a2, 0 ) - qe2 ), y2[ t ] = (signed char)( max( b2, 0 ) - qe2 );
                tbm[ ( r - t ) * MA_TINY_MAX + t ] = (unsigned char)d;
            }
        }
        // ksw_backtrack from (tlen - 1, qlen - 1); ops in backtrack order
        unsigned int cig[ 2 * MA_TINY_MAX + 2 ];
        int i = tlen - 1, j = qlen - 1, state = 0, n = 0;
        unsigned int cur = 0;
        while( i >= 0 && j >= 0 )
        {
            const unsigned int tmp = tbm[ j * MA_TINY_MAX + i ];
            if( state == 0 )
                state = tmp & 7;
            else if( !( tmp >> ( state + 2 ) & 1 ) )
                state = 0;
            if( state == 0 )
                state = tmp & 7;
            unsigned int op;
            if( state == 0 )
                op = 0, --i, --j;
            else if( state == 1 || state == 3 )
                op = 2, --i;
            else
                op = 1, --j;
            if( cur != 0 && ( cur & 0xf ) == op )
                cur += 1u << 4;
            else
            {
                if( cur != 0 )
                    cig[ n++ ] = cur;
                cur = 1u << 4 | op;
            }
        }
        if( i >= 0 )
        {
            if( cur != 0 && ( cur & 0xf ) == 2 )
                cur += (unsigned int)( i + 1 ) << 4;
            else
            {
                if( cur != 0 )
                    cig[ n++ ] = cur;
                cur = (unsigned int)( i + 1 ) << 4 | 2;
            }
        }
        if( j >= 0 )
        {
            if( cur != 0 && ( cur & 0xf ) == 1 )
                cur += (unsigned int)( j + 1 ) << 4;
            else
            {
                if( cur != 0 )
                    cig[ n++ ] = cur;
                cur = (unsigned int)( j + 1 ) << 4 | 1;
            }
        }
        if( cur != 0 )
            cig[ n++ ] = cur;
        KswOut ez;
        ez.max_q = ez.max_t = ez.mqe_t = ez.mte_q = -1;
        ez.max = 0;
        ez.score = ez.mqe = ez.mte = (int)0x80000000; // (not computed: the pipeline reads the CIGAR of a gap fill only)
        ez.zdropped = 0, ez.reach_end = 0, ez.status = 0;
        ez.cells = (long long)qlen * tlen;
        const unsigned long long o = n > 0 ? atomicAdd( A.cigar_cursor, (unsigned long long)n ) : 0ull;
        if( (long long)( o + n ) > A.cigar_cap )
        {
            atomicExch( A.error, 1 );
            ez.status = 1;
            n = 0;
        }
        const bool rev = T.flag & MA_KSW_REV_CIGAR;
        for( int k = 0; k < n; k++ )
            A.cigar[ o + k ] = rev ? cig[ k ] : cig[ n - 1 - k ];
        ez.n_cigar = n, ez.cigar_off = (long long)o;
        A.out[ ti ] = ez;
        cellsLocal += (unsigned long long)ez.cells;
    }
    if( A.cells_total && cellsLocal )
        atomicAdd( A.cells_total, cellsLocal );
}

} // namespace ma
