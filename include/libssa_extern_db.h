/*
 * libssa_extern_db.h - the database-provider interface the library imports
 * (lower boundary of the hot path; reference src/libssa_extern_db.h:20-47).
 *
 * libssa_b200.so ships a default FASTA provider for these four symbols as WEAK
 * definitions (libssa_b200/csrc/fasta_db.cpp), so a host program can link its
 * own provider (e.g. the reference author's libsdb) and override them.
 * IDs are 0-based record indices; ssa_db_get_sequence returns provider-owned
 * raw ASCII and NULL for id >= count (reference tests/test_libssa_extern_db.c:12-58).
 * The library pulls every record ONCE per init_db/init_symbol_translation to
 * build the device-resident packed database, and again per reported hit
 * (as reference src/algo/aligner.c:65 does).
 */
#ifndef LIBSSA_EXTERN_DB_H_
#define LIBSSA_EXTERN_DB_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    size_t ID;
    size_t seqlen;
    char * seq;
} seqinfo_t;
typedef seqinfo_t * p_seqinfo;

int ssa_db_init( const char * db_name );
size_t ssa_db_get_sequence_count( void );
p_seqinfo ssa_db_get_sequence( size_t id );
void ssa_db_close( void );

#ifdef __cplusplus
}
#endif

#endif /* LIBSSA_EXTERN_DB_H_ */
