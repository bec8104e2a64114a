// Host check: ma::stl::{sort,make_heap,pop_heap} perform exactly the element moves of libstdc++'s versions,
// including on tie-heavy input and on arrays that are not heaps.
#include "../../ma_b200/csrc/stl_exact.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <vector>
struct E { int key; int id; };
int main()
{
    srand( 12345 );
    auto cmp = []( const E& a, const E& b ) { return a.key < b.key; };
    long nTests = 0;
    for( int it = 0; it < 20000; it++ )
    {
        int n = it < 200 ? it : ( rand( ) % 3000 );
        int range = 1 + rand( ) % ( it % 3 == 0 ? 4 : it % 3 == 1 ? 50 : 100000 );
        std::vector<E> a( n );
        for( int i = 0; i < n; i++ ) a[ i ] = E{ rand( ) % range, i };
        if( it % 7 == 0 ) std::sort( a.begin( ), a.end( ), []( const E& x, const E& y ) { return x.key > y.key; } ); // adversarial-ish
        if( it % 11 == 0 && n > 40 ) for( int i = 0; i < n; i++ ) a[ i ].key = ( i * 7919 ) % 13; // organ-pipe like
        std::vector<E> b = a, c = a;
        std::sort( b.begin( ), b.end( ), cmp );
        ma::stl::sort( c.data( ), c.data( ) + n, cmp );
        for( int i = 0; i < n; i++ ) if( b[ i ].id != c[ i ].id ) { printf( "SORT MISMATCH n=%d it=%d at %d\n", n, it, i ); return 1; }
        // heap: make_heap, then rewrite keys (as rectangularSoC does) and pop everything
        b = a; c = a;
        std::make_heap( b.begin( ), b.end( ), cmp );
        ma::stl::make_heap( c.data( ), (long)n, cmp );
        for( int i = 0; i < n; i++ ) if( b[ i ].id != c[ i ].id ) { printf( "MAKE_HEAP MISMATCH\n" ); return 1; }
        for( int i = 0; i < n; i++ ) { int k = rand( ) % range; b[ i ].key = k; c[ i ].key = k; }
        for( long len = n; len > 0; len-- )
        {
            std::pop_heap( b.begin( ), b.begin( ) + len, cmp );
            ma::stl::pop_heap( c.data( ), len, cmp );
            for( int i = 0; i < n; i++ ) if( b[ i ].id != c[ i ].id ) { printf( "POP_HEAP MISMATCH n=%d len=%ld\n", n, len ); return 1; }
        }
        nTests++;
    }
    // force the heapsort fallback: median-of-3 killer sequence
    for( int n : { 1000, 4096, 10000 } )
    {
        std::vector<E> a( n );
        // Musser's killer adversary for median-of-3 quicksort
        int k = n / 2;
        for( int i = 0; i < k; i++ ) { if( i % 2 == 0 ) a[ i ] = E{ i + 1, i }; else a[ i ] = E{ k + i + ( k % 2 ? 0 : 1 ), i }; }
        for( int i = k; i < n; i++ ) a[ i ] = E{ ( i - k + 1 ) * 2, i };
        std::vector<E> b = a, c = a;
        std::sort( b.begin( ), b.end( ), cmp );
        ma::stl::sort( c.data( ), c.data( ) + n, cmp );
        for( int i = 0; i < n; i++ ) if( b[ i ].id != c[ i ].id ) { printf( "KILLER MISMATCH n=%d\n", n ); return 1; }
    }
    printf( "ok %ld\n", nTests );
    return 0;
}
