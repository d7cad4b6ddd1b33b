// Error/log conventions of the reference (src/util/util.c:34-98): fatal() prints to stderr and
// exits 1; info/warning/error go to stdout with a "libssa <LEVEL>: " prefix, gated by set_output_mode.
#include "common.h"

#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

int g_output_mode = OUTPUT_WARNING;  // reference src/util/util.c:34

void fatal( const char * format, ... ) {
    if( format ) {
        va_list ap;
        va_start( ap, format );
        vfprintf( stderr, format, ap );
        fprintf( stderr, "\n" );
        va_end( ap );
    }
    fflush( stdout );
    exit( 1 );
}

static void emit( int level, const char * tag, bool newline, const char * format, va_list ap ) {
    if( level > g_output_mode ) return;
    printf( "libssa %s: ", tag );
    vfprintf( stdout, format, ap );
    if( newline ) printf( "\n" );
}

void print_info( const char * format, ... ) {  // no trailing newline added (util.c:54-62)
    va_list ap;
    va_start( ap, format );
    emit( OUTPUT_INFO, "INFO", false, format, ap );
    va_end( ap );
}

void print_warning( const char * format, ... ) {
    va_list ap;
    va_start( ap, format );
    emit( OUTPUT_WARNING, "WARNING", true, format, ap );
    va_end( ap );
}

void print_error( const char * format, ... ) {
    va_list ap;
    va_start( ap, format );
    emit( OUTPUT_ERROR, "ERROR", true, format, ap );
    va_end( ap );
}

// 32-byte aligned like the reference (util.c:91-98) so memory handed to the caller can be released
// by either side's free().
void * xmalloc( size_t size ) {
    void * t = nullptr;
    if( posix_memalign( &t, 32, size ? size : 1 ) != 0 ) fatal( "Unable to allocate enough memory: %ld.", (long) size );
    return t;
}

void * xrealloc( void * ptr, size_t size ) {
    void * t = realloc( ptr, size ? size : 1 );
    if( !t ) fatal( "Unable to re-allocate enough memory: %ld.", (long) size );
    return t;
}

double wall_ms() {
    using namespace std::chrono;
    return duration<double, std::milli>( steady_clock::now().time_since_epoch() ).count();
}
