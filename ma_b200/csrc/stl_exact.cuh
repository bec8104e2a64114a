// Exact re-statement of the libstdc++ (GCC 13) algorithms whose tie / non-heap behaviour is OBSERVABLE in the
// reference's results: std::sort (introsort + final insertion sort), std::make_heap, std::pop_heap.
// The reference sorts seeds with many equal keys and pops from an array that is no longer a heap, so "any correct
// sort" is not enough — the element moves must be the same.  These templates perform the same comparisons and the
// same moves as bits/stl_algo.h / bits/stl_heap.h, on raw arrays, without recursion deeper than an explicit stack,
// and compile for host and device.
#pragma once

#if defined( __CUDACC__ )
#define MA_HD __host__ __device__
#define MA_NOINLINE __noinline__
#else
#define MA_HD
#define MA_NOINLINE __attribute__( ( noinline ) )
#endif

namespace ma
{
namespace stl
{

template <typename T> MA_HD inline void swp( T& a, T& b )
{
    T t = a;
    a = b;
    b = t;
}

// std::__push_heap
template <typename T, typename C> MA_HD inline void push_heap_( T* first, long hole, long top, T value, C comp )
{
    long parent = ( hole - 1 ) / 2;
    while( hole > top && comp( first[ parent ], value ) )
    {
        first[ hole ] = first[ parent ];
        hole = parent;
        parent = ( hole - 1 ) / 2;
    }
    first[ hole ] = value;
}

// std::__adjust_heap
template <typename T, typename C> MA_HD inline void adjust_heap( T* first, long hole, long len, T value, C comp )
{
    const long top = hole;
    long child = hole;
    while( child < ( len - 1 ) / 2 )
    {
        child = 2 * ( child + 1 );
        if( comp( first[ child ], first[ child - 1 ] ) )
            child--;
        first[ hole ] = first[ child ];
        hole = child;
    }
    if( ( len & 1 ) == 0 && child == ( len - 2 ) / 2 )
    {
        child = 2 * ( child + 1 );
        first[ hole ] = first[ child - 1 ];
        hole = child - 1;
    }
    push_heap_( first, hole, top, value, comp );
}

// std::make_heap
template <typename T, typename C> MA_HD inline void make_heap( T* first, long len, C comp )
{
    if( len < 2 )
        return;
    long parent = ( len - 2 ) / 2;
    while( true )
    {
        T value = first[ parent ];
        adjust_heap( first, parent, len, value, comp );
        if( parent == 0 )
            return;
        parent--;
    }
}

// std::pop_heap on [first, first+len): afterwards the former top sits at first[len-1]
template <typename T, typename C> MA_HD inline void pop_heap( T* first, long len, C comp )
{
    if( len > 1 )
    {
        T value = first[ len - 1 ];
        first[ len - 1 ] = first[ 0 ];
        adjust_heap( first, 0, len - 1, value, comp );
    }
}

template <typename T, typename C> MA_HD inline void unguarded_linear_insert( T* last, C comp )
{
    T val = *last;
    T* next = last - 1;
    while( comp( val, *next ) )
    {
        *last = *next;
        last = next;
        --next;
    }
    *last = val;
}

template <typename T, typename C> MA_HD inline void insertion_sort( T* first, T* last, C comp )
{
    if( first == last )
        return;
    for( T* i = first + 1; i != last; ++i )
    {
        if( comp( *i, *first ) )
        {
            T val = *i;
            for( T* p = i; p != first; --p )
                *p = *( p - 1 );
            *first = val;
        }
        else
            unguarded_linear_insert( i, comp );
    }
}

template <typename T, typename C> MA_HD inline void move_median_to_first( T* result, T* a, T* b, T* c, C comp )
{
    if( comp( *a, *b ) )
    {
        if( comp( *b, *c ) )
            swp( *result, *b );
        else if( comp( *a, *c ) )
            swp( *result, *c );
        else
            swp( *result, *a );
    }
    else if( comp( *a, *c ) )
        swp( *result, *a );
    else if( comp( *b, *c ) )
        swp( *result, *c );
    else
        swp( *result, *b );
}

template <typename T, typename C> MA_HD inline T* unguarded_partition( T* first, T* last, T* pivot, C comp )
{
    while( true )
    {
        while( comp( *first, *pivot ) )
            ++first;
        --last;
        while( comp( *pivot, *last ) )
            --last;
        if( !( first < last ) )
            return first;
        swp( *first, *last );
        ++first;
    }
}

// heapsort fallback: std::__partial_sort(first, last, last) == make_heap + sort_heap
template <typename T, typename C> MA_HD inline void heap_sort( T* first, T* last, C comp )
{
    make_heap( first, (long)( last - first ), comp );
    while( last - first > 1 )
    {
        --last;
        T value = *last;
        *last = *first;
        adjust_heap( first, 0, (long)( last - first ), value, comp );
    }
}

MA_HD inline int lg2( long n )
{
    int k = 0;
    while( n > 1 )
        n >>= 1, k++;
    return k;
}

// std::sort
template <typename T, typename C> MA_HD inline void sort( T* first, T* last, C comp )
{
    if( first == last )
        return;
    // __introsort_loop with the recursion on the right part turned into an explicit stack
    struct Frame
    {
        T *first, *last;
        int depth;
    };
    Frame stack[ 130 ];
    int sp = 0;
    stack[ sp++ ] = Frame{ first, last, lg2( last - first ) * 2 };
    while( sp > 0 )
    {
        Frame f = stack[ --sp ];
        while( f.last - f.first > 16 )
        {
            if( f.depth == 0 )
            {
                heap_sort( f.first, f.last, comp );
                break;
            }
            --f.depth;
            T* mid = f.first + ( f.last - f.first ) / 2;
            move_median_to_first( f.first, f.first + 1, mid, f.last - 1, comp );
            T* cut = unguarded_partition( f.first + 1, f.last, f.first, comp );
            // the reference recurses into [cut, last) first and then loops on [first, cut): the two ranges are
            // disjoint, so the order of processing does not change the result
            stack[ sp++ ] = Frame{ cut, f.last, f.depth };
            f.last = cut;
        }
    }
    // __final_insertion_sort
    if( last - first > 16 )
    {
        insertion_sort( first, first + 16, comp );
        for( T* i = first + 16; i != last; ++i )
            unguarded_linear_insert( i, comp );
    }
    else
        insertion_sort( first, last, comp );
}

} // namespace stl
} // namespace ma
