// Hit selection with the reference's exact tie behaviour.
//
// Which of several records tied at the cut-off score are reported depends, in the reference, on the
// ORDER in which scores reach a bounded min-heap that drops a newcomer equal to its current minimum
// (reference src/util/minheap.c:75-91, src/util/util.c:120-123).  With one thread that order is a
// deterministic function of the record lengths, the chunk size, the entry bit width and of which
// records overflow a narrow tier:
//   * chunks of `chunk` records in provider order, empty records skipped (src/db_adapter.c:212-239);
//     a chunk without any non-empty record ends the search (src/algo/8/search_8.c:135);
//   * 64-bit entry: provider order within the chunk (src/algo/64/search_64.c:44-56);
//   * 8/16-bit entry: the SIMD channel scheduler (src/algo/simd/search_simd_sw.c:338-428): 32 resp. 16
//     channels, each busy for ceil(len/4) blocks (fewer if the record overflows), finished channels
//     are emitted and refilled in channel order at the start of the next loop iteration;
//   * records that overflow are collected in emission order and re-run through the next tier after
//     the chunk (src/algo/8/search_8.c:94-124, src/algo/16/search_16.c:92-114).
// The GPU computes exact scores; this file only reproduces the order and the heap.
#include "common.h"

#include <algorithm>
#include <functional>

namespace {

// Emission order of one chunk through `channels` SIMD channels; blocks[i] = blocks record i occupies its channel.
// Event-driven form of the reference loop: a channel loaded at iteration T with a record of b blocks is found
// finished at the scan of iteration T + b; scans visit channels in index order, emitting and refilling as they go.
// A binary heap keyed by (finish iteration, channel) yields exactly that order in O(n log channels).
void simd_emission( const std::vector<uint64_t> & blocks, int channels, std::vector<uint32_t> & emit ) {
    const size_t n = blocks.size();
    emit.clear();
    emit.reserve( n );
    typedef std::pair<uint64_t, uint32_t> Event;  // (finish iteration << 6 | channel, record)
    std::vector<Event> heap;
    heap.reserve( channels );
    auto cmp = []( const Event & a, const Event & b ) { return a.first > b.first; };
    size_t next = 0;
    for( int c = 0; c < channels && next < n; c++ ) {  // first scan: all channels are empty and are filled in order
        heap.push_back( Event( ( blocks[next] << 6 ) | (uint64_t) c, (uint32_t) next ) );
        next++;
    }
    std::make_heap( heap.begin(), heap.end(), cmp );
    while( !heap.empty() ) {
        std::pop_heap( heap.begin(), heap.end(), cmp );
        const Event e = heap.back();
        heap.pop_back();
        emit.push_back( e.second );
        if( next < n ) {
            const uint64_t now = e.first >> 6, c = e.first & 63;
            heap.push_back( Event( ( ( now + blocks[next] ) << 6 ) | c, (uint32_t) next ) );
            std::push_heap( heap.begin(), heap.end(), cmp );
            next++;
        }
    }
}

template<typename Elem>
struct HeapT {  // reference src/util/minheap.c:30-91, comparisons on the score only
    size_t cap;
    std::vector<Elem> a;
    explicit HeapT( size_t k ) : cap( k ) { a.reserve( k ); }
    void add( const Elem & e ) {
        if( a.size() < cap ) {
            size_t i = a.size();
            a.push_back( e );
            while( i > 0 ) {
                size_t p = ( i - 1 ) / 2;
                if( !( e.score < a[p].score ) ) break;
                a[i] = a[p];
                i = p;
            }
            a[i] = e;
        }
        else if( cap > 0 && a[0].score < e.score ) {
            size_t p = 0, c = 1, n = a.size();
            while( c < n ) {
                if( c + 1 < n && a[c + 1].score < a[c].score ) c++;
                if( !( a[c].score < e.score ) ) break;
                a[p] = a[c];
                p = c;
                c = 2 * p + 1;
            }
            a[p] = e;
        }
    }
};
typedef HeapT<Hit> Heap;

int buffer_max_for_sort = 1;
int frame_hit_compare( const void * a, const void * b ) {  // minheap.c:93-102 / manager.c:68-77 on (score, record id), descending
    const FrameHit * x = (const FrameHit *) a;
    const FrameHit * y = (const FrameHit *) b;
    if( x->score != y->score ) return x->score < y->score ? 1 : -1;
    const uint32_t rx = x->slot / (uint32_t) buffer_max_for_sort, ry = y->slot / (uint32_t) buffer_max_for_sort;
    if( rx != ry ) return rx < ry ? 1 : -1;
    return 0;
}

}  // namespace

void arrival_order( const std::vector<uint64_t> & lens, const int64_t * ovf8, const int64_t * ovf16, size_t chunk,
        int bit_width, std::vector<uint32_t> & order ) {
    const size_t n = lens.size();
    order.clear();
    order.reserve( n );
    if( chunk == 0 ) chunk = DEFAULT_CHUNK_SIZE;
    std::vector<uint32_t> idx, emit, t16, t64;
    std::vector<uint64_t> blocks;
    for( size_t start = 0; start < n; start += chunk ) {
        idx.clear();
        for( size_t i = start; i < std::min( n, start + chunk ); i++ )
            if( lens[i] > 0 ) idx.push_back( (uint32_t) i );
        if( idx.empty() ) break;
        if( bit_width == 64 ) {
            order.insert( order.end(), idx.begin(), idx.end() );
            continue;
        }
        t16.clear();
        t64.clear();
        if( bit_width == 8 ) {
            blocks.resize( idx.size() );
            for( size_t i = 0; i < idx.size(); i++ ) {
                uint64_t b = ( lens[idx[i]] + 3 ) / 4;
                if( ovf8 && ovf8[idx[i]] >= 0 ) b = std::min<uint64_t>( b, (uint64_t) ovf8[idx[i]] + 1 );
                blocks[i] = b;
            }
            simd_emission( blocks, 32, emit );
            for( uint32_t e : emit ) {
                uint32_t s = idx[e];
                if( ovf8 && ovf8[s] >= 0 )
                    t16.push_back( s );
                else
                    order.push_back( s );
            }
        }
        else {
            t16 = idx;
        }
        if( !t16.empty() ) {
            blocks.resize( t16.size() );
            for( size_t i = 0; i < t16.size(); i++ ) {
                uint64_t b = ( lens[t16[i]] + 3 ) / 4;
                if( ovf16 && ovf16[t16[i]] >= 0 ) b = std::min<uint64_t>( b, (uint64_t) ovf16[t16[i]] + 1 );
                blocks[i] = b;
            }
            simd_emission( blocks, 16, emit );
            for( uint32_t e : emit ) {
                uint32_t s = t16[e];
                if( ovf16 && ovf16[s] >= 0 )
                    t64.push_back( s );
                else
                    order.push_back( s );
            }
        }
        order.insert( order.end(), t64.begin(), t64.end() );
    }
}

// manager.c:141-146: the thread heap is re-inserted in heap-array order, then both final sorts
static void finish_replay( const Heap & thread_heap, size_t k, std::vector<Hit> & out ) {
    Heap result( k );
    for( const Hit & h : thread_heap.a ) result.add( h );
    out = result.a;
    std::sort( out.begin(), out.end(), []( const Hit & x, const Hit & y ) {
        if( x.score != y.score ) return x.score > y.score;  // minheap.c:93-102, manager.c:68-77
        return x.db_id > y.db_id;
    } );
}

void replay_topk( const int32_t * scores, const uint32_t * order, size_t n, size_t k, std::vector<Hit> & out ) {
    Heap thread_heap( k );
    if( k > 0 ) {
        size_t i = 0;
        // fill phase
        for( ; i < n && thread_heap.a.size() < k; i++ ) thread_heap.add( Hit{ order[i], scores[i] } );
        // steady state: almost every record is rejected by one comparison with the root
        for( ; i < n; i++ )
            if( scores[i] > thread_heap.a[0].score ) thread_heap.add( Hit{ order[i], scores[i] } );
    }
    finish_replay( thread_heap, k, out );
}

// The same replay from the records that can touch the heap at all.  Once k records have arrived the root of the heap
// is the k-th largest score so far and never decreases, so a record that arrives after the first m (m >= k) and does not
// score above `threshold` = the root after those m is rejected whatever came in between: it need not be looked at.
// prefix_*: the first m arrivals; cand_*: the later arrivals scoring above the threshold, in arrival order.
// prefix_*: the first m arrivals; fetch( threshold, &scores, &ids ) supplies the later arrivals scoring above the threshold,
// in arrival order.
void replay_topk_sparse( const int32_t * prefix_scores, const uint32_t * prefix_ids, size_t m, size_t k,
        const std::function<size_t( int32_t, const int32_t **, const uint32_t ** )> & fetch, std::vector<Hit> & out ) {
    Heap thread_heap( k );
    if( k > 0 && m > 0 ) {
        size_t i = 0;
        for( ; i < m && thread_heap.a.size() < k; i++ ) thread_heap.add( Hit{ prefix_ids[i], prefix_scores[i] } );
        for( ; i < m; i++ )
            if( prefix_scores[i] > thread_heap.a[0].score ) thread_heap.add( Hit{ prefix_ids[i], prefix_scores[i] } );
        if( thread_heap.a.size() < k ) fatal( "libssa_b200: internal error: sparse replay with a prefix shorter than the hit list." );
        const int32_t * cs = nullptr;
        const uint32_t * ci = nullptr;
        const size_t ncand = fetch( (int32_t) thread_heap.a[0].score, &cs, &ci );
        for( size_t j = 0; j < ncand; j++ )
            if( cs[j] > thread_heap.a[0].score ) thread_heap.add( Hit{ ci[j], cs[j] } );
    }
    finish_replay( thread_heap, k, out );
}

// Searches with several query frames and/or several database entries per record.  Per chunk the reference runs every
// query frame over the chunk (src/algo/8/search_8.c:94-99), collects the entries that overflowed under ANY frame in
// detection order (an entry can be listed more than once) and re-runs that list under EVERY frame through the next
// tier (search_8.c:101-121, src/algo/16/search_16.c:92-113, src/algo/64/search_64.c:44-56), so a (frame, entry) pair
// can reach the heap more than once.  Entries of one record tie on (score, db_id) in both final sorts; the order among
// them is whatever the C library's qsort leaves, so the same qsort is applied here to the same array.
void general_search_order( const std::vector<uint64_t> & lens, size_t chunk_slots, int bit_width, int q_count, int buffer_max,
        const std::vector<std::vector<int32_t>> & scores, const std::vector<std::vector<int32_t>> & ovf8,
        const std::vector<std::vector<int32_t>> & ovf16, size_t hitcount, std::vector<FrameHit> & out, size_t * n_ovf8,
        size_t * n_ovf16 ) {
    size_t count8 = 0, count16 = 0;
    const size_t n = lens.size();
    HeapT<FrameHit> thread_heap( hitcount );
    auto add = [&]( int q, uint32_t slot ) { thread_heap.add( FrameHit{ slot, (uint8_t) q, (int64_t) scores[q][slot] } ); };
    auto overflows = []( const std::vector<std::vector<int32_t>> & o, int q, uint32_t slot ) {
        return !o.empty() && !o[q].empty() && o[q][slot] >= 0;
    };
    std::vector<uint32_t> idx, emit, l16, l64;
    std::vector<uint64_t> blocks;
    // one SIMD tier over `list` for every frame: non-overflowing entries reach the heap, the others are appended to `next`
    auto simd_tier = [&]( const std::vector<uint32_t> & list, int channels, const std::vector<std::vector<int32_t>> & ovf,
                          std::vector<uint32_t> & next ) {
        for( int q = 0; q < q_count; q++ ) {
            blocks.resize( list.size() );
            for( size_t i = 0; i < list.size(); i++ ) {
                uint64_t b = ( lens[list[i]] + 3 ) / 4;
                if( overflows( ovf, q, list[i] ) ) b = std::min<uint64_t>( b, (uint64_t) ovf[q][list[i]] + 1 );
                blocks[i] = b;
            }
            simd_emission( blocks, channels, emit );
            for( uint32_t e : emit ) {
                const uint32_t s = list[e];
                if( overflows( ovf, q, s ) )
                    next.push_back( s );
                else
                    add( q, s );
            }
        }
    };
    for( size_t start = 0; start < n; start += chunk_slots ) {
        idx.clear();
        for( size_t i = start; i < std::min( n, start + chunk_slots ); i++ )
            if( lens[i] > 0 ) idx.push_back( (uint32_t) i );
        if( idx.empty() ) break;
        l16.clear();
        l64.clear();
        if( bit_width == 64 ) {
            l64 = idx;
        }
        else {
            if( bit_width == 8 )
                simd_tier( idx, 32, ovf8, l16 );
            else
                l16 = idx;
            if( bit_width == 8 ) count8 += l16.size();  // search_8.c:119
            if( !l16.empty() ) simd_tier( l16, 16, ovf16, l64 );
            count16 += l64.size();  // search_16.c:106
        }
        for( int q = 0; q < q_count && !l64.empty(); q++ )
            for( uint32_t s : l64 ) add( q, s );
    }
    if( n_ovf8 ) *n_ovf8 = count8;
    if( n_ovf16 ) *n_ovf16 = count16;
    HeapT<FrameHit> result( hitcount );  // manager.c:141-146
    for( const FrameHit & h : thread_heap.a ) result.add( h );
    out = result.a;
    buffer_max_for_sort = buffer_max;
    if( !out.empty() ) {
        qsort( out.data(), out.size(), sizeof( FrameHit ), frame_hit_compare );  // minheap_sort, manager.c:176
        qsort( out.data(), out.size(), sizeof( FrameHit ), frame_hit_compare );  // sort_alignment_list, manager.c:113
    }
}
