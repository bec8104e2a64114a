// Shared host-side plumbing for libma_b200.so: error handling, grow-only device buffers, the context.
#pragma once
#include "../../include/ma_b200.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace ma
{

struct CudaError
{
    std::string msg;
};

#define MA_CUDA( call )                                                                                                \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = ( call );                                                                                     \
        if( e_ != cudaSuccess )                                                                                        \
            throw ma::CudaError{ std::string( #call ) + ": " + cudaGetErrorString( e_ ) + " (" + __FILE__ + ":" +      \
                                 std::to_string( __LINE__ ) + ")" };                                                   \
    } while( 0 )

// grow-only device buffer; contents are NOT preserved across a growing reserve()
template <typename T> struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0;
    void reserve( size_t n )
    {
        if( n <= cap )
            return;
        if( p )
            cudaFree( p );
        p = nullptr;
        cap = 0;
        size_t want = n + n / 8 + 64;
        MA_CUDA( cudaMalloc( (void**)&p, want * sizeof( T ) ) );
        cap = want;
    }
    void release( )
    {
        if( p )
            cudaFree( p );
        p = nullptr;
        cap = 0;
    }
    ~DevBuf( )
    {
        release( );
    }
};

struct EventTimer
{
    cudaEvent_t a = nullptr, b = nullptr;
    void init( )
    {
        MA_CUDA( cudaEventCreate( &a ) );
        MA_CUDA( cudaEventCreate( &b ) );
    }
    void start( cudaStream_t s )
    {
        MA_CUDA( cudaEventRecord( a, s ) );
    }
    float stop( cudaStream_t s )
    {
        MA_CUDA( cudaEventRecord( b, s ) );
        MA_CUDA( cudaEventSynchronize( b ) );
        float ms = 0;
        MA_CUDA( cudaEventElapsedTime( &ms, a, b ) );
        return ms;
    }
};

} // namespace ma
