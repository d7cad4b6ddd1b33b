/*
 * TEST INFRASTRUCTURE ONLY - never linked into, imported by or called from the
 * product (libssa_b200/).  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may use it.
 *
 * Plain-C restatement of the reference's algorithm for the hot path
 * (score every database sequence, keep the top k, trace back the reported
 * hits).  Each function cites the reference file:line it follows.  Parity is
 * PINNED: tests/test_oracle.py checks this file against the compiled,
 * unmodified reference (oracle/_ref/libssa_ref.so) and against the golden
 * vectors of the reference's own tests.
 *
 * All sequences are code strings (0..31), the score matrix is the reference's
 * 32x32 table indexed (db_symbol << 5) + query_symbol unless stated otherwise
 * (reference src/matrices.h:26-44).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define MAT(M, a, b) ((M)[(((int)(a)) << 5) + (int)(b)])

/* ------------------------------------------------------------------------- */
/* Scores: reference src/algo/64/smith_waterman_63.c:32-98 and               */
/*         src/algo/64/needleman_wunsch_64.c:32-98                            */
/* ------------------------------------------------------------------------- */

int64_t orc_sw_score( const uint8_t * d, size_t dlen, const uint8_t * q, size_t qlen, const int64_t * M, int gapO,
        int gapE ) {
    int64_t * H = calloc( qlen ? qlen : 1, sizeof(int64_t) ); /* H of the previous column, per query row */
    int64_t * E = calloc( qlen ? qlen : 1, sizeof(int64_t) ); /* E entering this column, per query row */
    int64_t best = 0;
    for( size_t j = 0; j < dlen; j++ ) {
        int64_t diag = 0, f = 0;
        for( size_t i = 0; i < qlen; i++ ) {
            int64_t up_left = H[i];
            int64_t h = diag + MAT( M, d[j], q[i] );
            if( E[i] > h ) h = E[i];
            if( f > h ) h = f;
            if( h < 0 ) h = 0;
            if( h > best ) best = h;
            H[i] = h;
            int64_t open = h + gapO + gapE;
            E[i] = ( E[i] + gapE > open ) ? E[i] + gapE : open;
            f = ( f + gapE > open ) ? f + gapE : open;
            diag = up_left;
        }
    }
    free( H );
    free( E );
    return best;
}

int64_t orc_nw_score( const uint8_t * d, size_t dlen, const uint8_t * q, size_t qlen, const int64_t * M, int gapO,
        int gapE ) {
    int64_t * H = malloc( ( qlen ? qlen : 1 ) * sizeof(int64_t) );
    int64_t * E = malloc( ( qlen ? qlen : 1 ) * sizeof(int64_t) );
    for( size_t i = 0; i < qlen; i++ ) {
        H[i] = gapO + (int64_t)( i + 1 ) * gapE;         /* needleman_wunsch_64.c:40 */
        E[i] = 2 * (int64_t) gapO + (int64_t)( i + 2 ) * gapE; /* :41 */
    }
    for( size_t j = 0; j < dlen; j++ ) {
        int64_t f = 2 * (int64_t) gapO + (int64_t)( j + 2 ) * gapE;    /* :51 */
        int64_t diag = ( j == 0 ) ? 0 : gapO + (int64_t) j * gapE; /* :52 */
        for( size_t i = 0; i < qlen; i++ ) {
            int64_t up_left = H[i];
            int64_t h = diag + MAT( M, d[j], q[i] );
            if( f > h ) h = f;
            if( E[i] > h ) h = E[i];
            H[i] = h;
            int64_t open = h + gapO + gapE;
            E[i] = ( E[i] + gapE > open ) ? E[i] + gapE : open;
            f = ( f + gapE > open ) ? f + gapE : open;
            diag = up_left;
        }
    }
    int64_t r = qlen ? H[qlen - 1] : 0;
    free( H );
    free( E );
    return r;
}

/*
 * SW score plus, for the emulation of the narrow SIMD tiers, the index of the
 * first 4-column block in which the running maximum reaches `thr8` (255) and
 * `thr16` (65535): reference src/algo/simd/search_simd_sw.c:422-423 forces a
 * channel change after the block in which S saturates.  -1 = never.
 */
int64_t orc_sw_score_blocks( const uint8_t * d, size_t dlen, const uint8_t * q, size_t qlen, const int64_t * M,
        int gapO, int gapE, int64_t thr8, int64_t thr16, int64_t * block8, int64_t * block16 ) {
    int64_t * H = calloc( qlen ? qlen : 1, sizeof(int64_t) );
    int64_t * E = calloc( qlen ? qlen : 1, sizeof(int64_t) );
    int64_t best = 0;
    *block8 = -1;
    *block16 = -1;
    for( size_t j = 0; j < dlen; j++ ) {
        int64_t diag = 0, f = 0;
        for( size_t i = 0; i < qlen; i++ ) {
            int64_t up_left = H[i];
            int64_t h = diag + MAT( M, d[j], q[i] );
            if( E[i] > h ) h = E[i];
            if( f > h ) h = f;
            if( h < 0 ) h = 0;
            if( h > best ) best = h;
            H[i] = h;
            int64_t open = h + gapO + gapE;
            E[i] = ( E[i] + gapE > open ) ? E[i] + gapE : open;
            f = ( f + gapE > open ) ? f + gapE : open;
            diag = up_left;
        }
        if( *block8 < 0 && best >= thr8 ) *block8 = (int64_t)( j / 4 );
        if( *block16 < 0 && best >= thr16 ) *block16 = (int64_t)( j / 4 );
    }
    free( H );
    free( E );
    return best;
}

/*
 * Scalar emulation of ONE channel of the 8-/16-bit saturating SIMD NW kernel,
 * reference src/algo/simd/search_simd_nw.c:180-316 (ALIGNCORE,
 * aligncolumns_first/rest) and :390-500 (boundary injection, per-block
 * overflow test, result read-out).  A channel's arithmetic does not depend on
 * its neighbours, so this reproduces exactly what the reference computes for
 * one database sequence: the narrow-tier score, or the 4-column block after
 * which the channel is abandoned because of overflow (*ovf_block >= 0; the
 * sequence is then re-run in the next tier).  bits = 8 or 16.
 */
static inline int64_t sat( int64_t v, int64_t lo, int64_t hi ) {
    return v < lo ? lo : ( v > hi ? hi : v );
}

static inline int64_t wrap( int64_t v, int bits ) {
    return bits == 8 ? (int64_t)(int8_t) v : (int64_t)(int16_t) v;
}

int64_t orc_nw_simd_channel( const uint8_t * d, size_t dlen, const uint8_t * q, size_t qlen, const int64_t * M,
        int gapO, int gapE, int bits, int64_t * ovf_block ) {
    const int64_t LO = bits == 8 ? INT8_MIN : INT16_MIN, HI = bits == 8 ? INT8_MAX : INT16_MAX;
    const int64_t QR = wrap( gapO + gapE, bits ), R = wrap( gapE, bits );
    const int64_t score_min = wrap( LO - gapO - gapE - 1, bits ); /* search_simd_nw.c:372 */
    int64_t * H = calloc( qlen ? qlen : 1, sizeof(int64_t) );
    int64_t * E = calloc( qlen ? qlen : 1, sizeof(int64_t) );
    int64_t Hb[4], Fb[4], Sm[4] = { 0, 0, 0, 0 };
    Hb[0] = 0;
    for( int k = 1; k < 4; k++ ) Hb[k] = wrap( gapO + k * gapE, bits );     /* :445-448 */
    for( int k = 0; k < 4; k++ ) Fb[k] = wrap( gapO + ( k + 1 ) * gapE, bits ); /* :450-453 */
    size_t nblocks = ( dlen + 3 ) / 4;
    *ovf_block = -1;
    for( size_t blk = 0; blk < nblocks; blk++ ) {
        uint8_t r[4];
        for( int k = 0; k < 4; k++ ) r[k] = ( 4 * blk + k < dlen ) ? d[4 * blk + k] : 0;
        int64_t h[4], f[4], hn[4];
        for( int k = 0; k < 4; k++ ) {
            h[k] = Hb[k];
            f[k] = sat( Fb[k] + QR, LO, HI );
        }
        int64_t hmin = 0, hmax = 0, mge = QR;
        for( size_t i = 0; i < qlen; i++ ) {
            int64_t h4, e;
            if( blk == 0 ) { /* masked initialisation of a restarting channel, :205-240 */
                h4 = sat( 0 + mge, LO, HI );
                e = sat( sat( 0 + mge, LO, HI ) + QR, LO, HI );
                mge = sat( mge + R, LO, HI );
            }
            else {
                h4 = H[i];
                e = E[i];
            }
            for( int k = 0; k < 4; k++ ) {
                int64_t v = wrap( MAT( M, r[k], q[i] ), bits );
                int64_t x = sat( h[k] + v, LO, HI );
                if( f[k] > x ) x = f[k];
                if( e > x ) x = e;
                if( x < hmin ) hmin = x;
                if( x > hmax ) hmax = x;
                hn[k] = x;
                x = sat( x + QR, LO, HI );
                f[k] = sat( f[k] + R, LO, HI );
                if( x > f[k] ) f[k] = x;
                e = sat( e + R, LO, HI );
                if( x > e ) e = x;
            }
            H[i] = hn[3];
            E[i] = e;
            h[0] = h4;
            h[1] = hn[0];
            h[2] = hn[1];
            h[3] = hn[2];
        }
        Sm[0] = h[1];
        Sm[1] = h[2];
        Sm[2] = h[3];
        Sm[3] = qlen ? H[qlen - 1] : 0;
        if( score_min > hmin || hmax == HI ) { /* :483-485 */
            *ovf_block = (int64_t) blk;
            break;
        }
        Fb[0] = sat( Fb[3] + R, LO, HI ); /* :494-502 */
        for( int k = 1; k < 4; k++ ) Fb[k] = sat( Fb[k - 1] + R, LO, HI );
        Hb[0] = sat( Hb[3] + R, LO, HI );
        for( int k = 1; k < 4; k++ ) Hb[k] = sat( Hb[k - 1] + R, LO, HI );
    }
    int64_t score = Sm[( dlen + 3 ) % 4]; /* :423-424 */
    if( *ovf_block < 0 && !( score > LO && score < HI ) ) /* :426 */
        *ovf_block = nblocks ? (int64_t) nblocks - 1 : 0;
    free( H );
    free( E );
    return score;
}

/* scores of a whole database (concatenated code strings, offsets[n+1]) */
void orc_score_db( int nw, const uint8_t * db, const uint64_t * offsets, size_t n, const uint8_t * q, size_t qlen,
        const int64_t * M, int gapO, int gapE, int64_t * out ) {
    for( size_t s = 0; s < n; s++ ) {
        const uint8_t * d = db + offsets[s];
        size_t dlen = offsets[s + 1] - offsets[s];
        out[s] = nw ? orc_nw_score( d, dlen, q, qlen, M, gapO, gapE ) : orc_sw_score( d, dlen, q, qlen, M, gapO, gapE );
    }
}

/* ------------------------------------------------------------------------- */
/* Bounded top-k: reference src/util/minheap.c:30-106, src/util/util.c:108-124 */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint64_t id;
    int64_t score;
} orc_elem;

typedef struct {
    size_t cap, count;
    orc_elem * a;
} orc_heap;

static void heap_add( orc_heap * h, orc_elem e ) {
    if( h->count < h->cap ) {
        /* minheap.c:76-86: append and sift up while strictly smaller than the parent */
        size_t i = h->count++;
        while( i > 0 ) {
            size_t p = ( i - 1 ) / 2;
            if( !( e.score < h->a[p].score ) ) break;
            h->a[i] = h->a[p];
            i = p;
        }
        h->a[i] = e;
    }
    else if( h->cap > 0 && h->a[0].score < e.score ) {
        /* minheap.c:87-90 + :52-73: replace root only if strictly greater, sift down */
        size_t p = 0, c = 1;
        while( c < h->count ) {
            if( c + 1 < h->count && h->a[c + 1].score < h->a[c].score ) c++;
            if( h->a[c].score < e.score )
                h->a[p] = h->a[c];
            else
                break;
            p = c;
            c = 2 * p + 1;
        }
        h->a[p] = e;
    }
}

static int cmp_desc( const void * x, const void * y ) {
    /* minheap.c:93-102 / manager.c:68-77: score descending, then id descending */
    const orc_elem * a = x, * b = y;
    if( a->score != b->score ) return a->score > b->score ? -1 : 1;
    if( a->id != b->id ) return a->id > b->id ? -1 : 1;
    return 0;
}

/*
 * Replays a one-thread search: elements arrive in `order` (indices into
 * scores/ids) at the thread's heap (searcher.c:155-174), the thread heap is
 * re-inserted in heap-array order into the result heap (manager.c:141-146),
 * which is then sorted (manager.c:172).  Returns the number of hits written.
 */
size_t orc_topk_replay( const int64_t * scores, const uint64_t * ids, const uint64_t * order, size_t n, size_t k,
        uint64_t * out_ids, int64_t * out_scores ) {
    orc_heap th = { k, 0, malloc( ( k ? k : 1 ) * sizeof(orc_elem) ) };
    orc_heap res = { k, 0, malloc( ( k ? k : 1 ) * sizeof(orc_elem) ) };
    for( size_t i = 0; i < n; i++ ) {
        size_t s = order ? order[i] : i;
        orc_elem e = { ids ? ids[s] : s, scores[s] };
        heap_add( &th, e );
    }
    for( size_t i = 0; i < th.count; i++ )
        heap_add( &res, th.a[i] );
    qsort( res.a, res.count, sizeof(orc_elem), cmp_desc );
    for( size_t i = 0; i < res.count; i++ ) {
        out_ids[i] = res.a[i].id;
        out_scores[i] = res.a[i].score;
    }
    size_t r = res.count;
    free( th.a );
    free( res.a );
    return r;
}

/* ------------------------------------------------------------------------- */
/* Arrival order of a one-thread search                                       */
/* ------------------------------------------------------------------------- */

/*
 * Channel scheduler of the SIMD kernels, reference
 * src/algo/simd/search_simd_sw.c:338-428 (NW: search_simd_nw.c:390-500),
 * driven by lengths only.  Every loop iteration advances each busy channel by
 * one 4-column block; at the start of an iteration that follows a block in
 * which some channel ran dry (or overflowed) the channels are scanned in index
 * order, each finished one emits its sequence and is refilled with the next
 * sequence of the chunk.  blocks[i] = number of blocks sequence i occupies its
 * channel (ceil(len/4), or fewer when it overflows early).
 * Writes the emission order (indices 0..n-1 into the chunk) to `emit`.
 */
static void simd_emission_order( const uint64_t * blocks, size_t n, int channels, uint64_t * emit ) {
    int64_t * cur = malloc( channels * sizeof(int64_t) );  /* sequence in channel, -1 = empty */
    uint64_t * left = malloc( channels * sizeof(uint64_t) ); /* blocks still to run */
    for( int c = 0; c < channels; c++ ) {
        cur[c] = -1;
        left[c] = 0;
    }
    size_t next = 0, done = 0, ne = 0;
    while( done < n ) {
        /* scan: emit + refill the channels that have nothing left */
        for( int c = 0; c < channels; c++ ) {
            if( left[c] == 0 ) {
                if( cur[c] >= 0 ) {
                    emit[ne++] = (uint64_t) cur[c];
                    done++;
                    cur[c] = -1;
                }
                if( next < n ) {
                    cur[c] = (int64_t) next;
                    left[c] = blocks[next];
                    next++;
                }
            }
        }
        if( done == n ) break;
        /* run blocks until some channel runs dry */
        uint64_t step = UINT64_MAX;
        for( int c = 0; c < channels; c++ )
            if( cur[c] >= 0 && left[c] < step ) step = left[c];
        for( int c = 0; c < channels; c++ )
            if( cur[c] >= 0 ) left[c] -= step;
    }
    free( cur );
    free( left );
}

/*
 * Full one-thread arrival order for a search started at `bit_width`
 * (8, 16 or 64) with `chunk` sequences per chunk (db_adapter.c:212-239,
 * search_8.c:94-146, search_16.c:92-134, search_64.c:44-79).
 *
 * lens[n]: sequence lengths in database order (zero-length records are
 *          skipped, db_adapter.c:231-233; a chunk with no non-empty record
 *          ends the search, search_*.c `while( chunk->fill_pointer )`).
 * ovf8[n]/ovf16[n]: -1, or the 0-based 4-column block in which the sequence
 *          overflows the 8-/16-bit tier (it then leaves its channel after that
 *          block and is re-run in the next tier after the chunk).
 * order[]: receives indices into lens[] in heap-arrival order; returns count.
 */
size_t orc_arrival_order( const uint64_t * lens, const int64_t * ovf8, const int64_t * ovf16, size_t n, size_t chunk,
        int bit_width, uint64_t * order ) {
    size_t out = 0;
    uint64_t * idx = malloc( ( chunk ? chunk : 1 ) * sizeof(uint64_t) );
    uint64_t * blocks = malloc( ( chunk ? chunk : 1 ) * sizeof(uint64_t) );
    uint64_t * emit = malloc( ( chunk ? chunk : 1 ) * sizeof(uint64_t) );
    uint64_t * o8 = malloc( ( chunk ? chunk : 1 ) * sizeof(uint64_t) );
    uint64_t * o16 = malloc( ( chunk ? chunk : 1 ) * sizeof(uint64_t) );
    for( size_t start = 0; start < n; start += chunk ) {
        size_t m = 0;
        for( size_t i = start; i < start + chunk && i < n; i++ )
            if( lens[i] > 0 ) idx[m++] = i;
        if( m == 0 ) break;
        if( bit_width == 64 ) {
            for( size_t i = 0; i < m; i++ ) order[out++] = idx[i];
            continue;
        }
        size_t n16 = 0, n8 = 0;
        if( bit_width == 8 ) {
            for( size_t i = 0; i < m; i++ ) {
                uint64_t b = ( lens[idx[i]] + 3 ) / 4;
                if( ovf8 && ovf8[idx[i]] >= 0 && (uint64_t) ovf8[idx[i]] + 1 < b ) b = (uint64_t) ovf8[idx[i]] + 1;
                blocks[i] = b;
            }
            simd_emission_order( blocks, m, 32, emit );
            for( size_t i = 0; i < m; i++ ) {
                uint64_t s = idx[emit[i]];
                if( ovf8 && ovf8[s] >= 0 )
                    o8[n8++] = s;
                else
                    order[out++] = s;
            }
        }
        else {
            for( size_t i = 0; i < m; i++ ) o8[n8++] = idx[i];
        }
        /* 16-bit tier over the chunk (or over the 8-bit overflow chunk) */
        if( n8 > 0 ) {
            for( size_t i = 0; i < n8; i++ ) {
                uint64_t b = ( lens[o8[i]] + 3 ) / 4;
                if( ovf16 && ovf16[o8[i]] >= 0 && (uint64_t) ovf16[o8[i]] + 1 < b ) b = (uint64_t) ovf16[o8[i]] + 1;
                blocks[i] = b;
            }
            simd_emission_order( blocks, n8, 16, emit );
            for( size_t i = 0; i < n8; i++ ) {
                uint64_t s = o8[emit[i]];
                if( ovf16 && ovf16[s] >= 0 )
                    o16[n16++] = s;
                else
                    order[out++] = s;
            }
        }
        /* 64-bit tier: overflow chunk in the order it was filled */
        for( size_t i = 0; i < n16; i++ ) order[out++] = o16[i];
    }
    free( idx );
    free( blocks );
    free( emit );
    free( o8 );
    free( o16 );
    return out;
}

/* ------------------------------------------------------------------------- */
/* Alignment of one reported hit                                              */
/* ------------------------------------------------------------------------- */

/*
 * Local-alignment region: reference src/algo/align.c:39-139.  a = query,
 * b = database sequence.  The matrix is read as M[a][b] here (align.c:64,106).
 * The reverse pass re-initialises HH/EE only for indices 0..b_end although
 * they are indexed by the query position (align.c:86-89); entries beyond keep
 * the values of the forward pass.  Returns 0 on success, -1 for the
 * reference's "Internal error in align function" (align.c:131).
 */
int orc_region_local( const uint8_t * a, size_t alen, const uint8_t * b, size_t blen, const int64_t * M, int gapO,
        int gapE, int64_t * region /* a_begin,a_end,b_begin,b_end */) {
    size_t size = alen > blen ? alen : blen;
    int64_t * HH = malloc( ( size ? size : 1 ) * sizeof(int64_t) );
    int64_t * EE = malloc( ( size ? size : 1 ) * sizeof(int64_t) );
    int64_t score = 0;
    int64_t a_end = 0, b_end = 0, a_begin = 0, b_begin = 0;

    for( size_t j = 0; j < alen; j++ ) {
        HH[j] = 0;
        EE[j] = gapO;
    }
    for( size_t i = 0; i < blen; i++ ) {
        int64_t h = 0, p = 0, f = gapO;
        for( size_t j = 0; j < alen; j++ ) {
            f = ( f > h + gapO ? f : h + gapO ) + gapE;
            EE[j] = ( EE[j] > HH[j] + gapO ? EE[j] : HH[j] + gapO ) + gapE;
            h = p + MAT( M, a[j], b[i] );
            if( h < 0 ) h = 0;
            if( f > h ) h = f;
            if( EE[j] > h ) h = EE[j];
            p = HH[j];
            HH[j] = h;
            if( h > score ) {
                score = h;
                a_end = (int64_t) j;
                b_end = (int64_t) i;
            }
        }
    }

    for( int64_t j = b_end; j >= 0; j-- ) {
        if( (size_t) j < size ) {
            HH[j] = -1;
            EE[j] = -1;
        }
    }
    int64_t cost = 0;
    int found = 0;
    for( int64_t i = b_end; i >= 0 && !found; i-- ) {
        int64_t h = -1, f = -1;
        int64_t p = ( i == b_end ) ? 0 : -1;
        for( int64_t j = a_end; j >= 0; j-- ) {
            f = ( f > h + gapO ? f : h + gapO ) + gapE;
            EE[j] = ( EE[j] > HH[j] + gapO ? EE[j] : HH[j] + gapO ) + gapE;
            h = p + MAT( M, a[j], b[i] );
            if( f > h ) h = f;
            if( EE[j] > h ) h = EE[j];
            p = HH[j];
            HH[j] = h;
            if( h > cost ) {
                cost = h;
                a_begin = j;
                b_begin = i;
                if( cost >= score ) {
                    found = 1;
                    break;
                }
            }
        }
    }
    free( HH );
    free( EE );
    region[0] = a_begin;
    region[1] = a_end;
    region[2] = b_begin;
    region[3] = b_end;
    return found ? 0 : -1;
}

#define DIR_UP 1
#define DIR_LEFT 2
#define DIR_EXT_UP 4
#define DIR_EXT_LEFT 8

/*
 * Direction matrix + backward walk + run-length string: reference
 * src/algo/cigar.c:48-198 (directions), :299-335 (walk), :221-263 (string).
 * Matrix read as M[b][a] (cigar.c:79,152).  Returns a malloc'ed string.
 */
char * orc_cigar( int nw, const uint8_t * a, size_t alen, const uint8_t * b, size_t blen, const int64_t * M, int gapO,
        int gapE, const int64_t * region ) {
    uint8_t * dir = calloc( ( alen && blen ) ? alen * blen : 1, 1 );
    int64_t * H = malloc( ( alen ? alen : 1 ) * sizeof(int64_t) );
    int64_t * E = malloc( ( alen ? alen : 1 ) * sizeof(int64_t) );
    for( size_t i = 0; i < alen; i++ ) {
        H[i] = nw ? gapO + (int64_t)( i + 1 ) * gapE : 0;
        E[i] = nw ? 2 * (int64_t) gapO + (int64_t)( i + 2 ) * gapE : 0;
    }
    for( size_t j = 0; j < blen; j++ ) {
        int64_t f = nw ? 2 * (int64_t) gapO + (int64_t)( j + 2 ) * gapE : 0;
        int64_t diag = nw ? ( j == 0 ? 0 : gapO + (int64_t) j * gapE ) : 0;
        for( size_t i = 0; i < alen; i++ ) {
            uint8_t bits = 0;
            int64_t up_left = H[i];
            int64_t e = E[i];
            int64_t h = diag + MAT( M, b[j], a[i] );
            if( f > h ) {
                bits |= DIR_UP;
                h = f;
            }
            if( e > h ) {
                bits |= DIR_LEFT;
                h = e;
            }
            if( !nw && h < 0 ) h = 0;
            H[i] = h;
            h += gapO + gapE;
            e += gapE;
            f += gapE;
            if( f > h )
                bits |= DIR_EXT_UP;
            else
                f = h;
            if( e > h )
                bits |= DIR_EXT_LEFT;
            else
                e = h;
            E[i] = e;
            dir[alen * j + i] = bits;
            diag = up_left;
        }
    }
    free( H );
    free( E );

    /* walk */
    size_t cap = 64, len = 0;
    char * ops = malloc( cap ); /* one char per step, in walk order */
    int64_t i = region[1], j = region[3];
    while( i >= 0 && j >= 0 && i >= region[0] && j >= region[2] ) {
        uint8_t d = dir[alen * (size_t) j + (size_t) i];
        char op;
        if( d & ( DIR_LEFT | DIR_EXT_LEFT ) ) {
            j--;
            op = 'I';
        }
        else if( d & ( DIR_UP | DIR_EXT_UP ) ) {
            i--;
            op = 'D';
        }
        else {
            i--;
            j--;
            op = 'M';
        }
        if( len == cap ) {
            cap *= 2;
            ops = realloc( ops, cap );
        }
        ops[len++] = op;
    }
    free( dir );

    /* run-length encode front to back of the alignment (= reverse walk order); count omitted when 1 */
    size_t ocap = 2 * len + 32, olen = 0;
    char * out = malloc( ocap );
    size_t k = len;
    while( k > 0 ) {
        char op = ops[k - 1];
        size_t run = 0;
        while( k > 0 && ops[k - 1] == op ) {
            run++;
            k--;
        }
        if( olen + 32 > ocap ) {
            ocap = 2 * ocap + 32;
            out = realloc( out, ocap );
        }
        if( run > 1 ) olen += (size_t) sprintf( out + olen, "%zu", run );
        out[olen++] = op;
    }
    out[olen] = 0;
    free( ops );
    return out;
}

void orc_free( void * p ) {
    free( p );
}
