/* kaarme_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded CPU restatement of the counting hot path of the reference
 * (Denopia/canonical-k-mer-hash-table, "Kaarme").  It exists so that tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs can CHECK the CUDA path.  Nothing on the
 * product path (canonical-k-mer-hash-table_b200/csrc, host/) includes, links or executes this file.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 *   - golden vectors minted from the reference's own code (tests/golden/, made by
 *     tests/golden/make_golden.py running oracle/_ref/kaarme and oracle/_ref/xxh64_kat), and
 *   - when oracle/_ref/kaarme is present, the reference binary itself on freshly seeded inputs.
 *
 * Every function cites the reference file:line it restates (paths relative to /root/reference).
 */
#ifndef KAARME_ORACLE_H
#define KAARME_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KO_FASTA 0 /* main.cpp:179-181 input_mode 0 */
#define KO_PLAIN 2 /* main.cpp:185-187 input_mode 2 */

#define KO_TABLE_PLAIN 0  /* -m 0: uint16 counter that wraps  (kmer_hash_table.hpp:42)          */
#define KO_TABLE_KAARME 2 /* -m 2: 14-bit counter that saturates at 16383 (kmer.cpp:699-714)      */
#define KO_TABLE_EXACT (-1) /* no clamp: the true multiplicity                                      */

/* ---- scalar helpers ------------------------------------------------------------------------- */

/* functions_strings.cpp:56-70 : A/a C/c G/g T/t -> 0..3, anything else -> 4 */
uint32_t ko_char2int(uint8_t c);

/* functions_math.cpp:53-96 : smallest prime p >= at_least with p % 4 == 3 (2 if at_least <= 2) */
uint64_t ko_next_prime3mod4(uint64_t at_least);

/* functions_math.cpp:99-130 : modular inverse of A mod M (extended Euclid, A and M coprime) */
uint64_t ko_modinv(int64_t A, int64_t M);

/* external/xxHash/xxhash.h:3495 (XXH64) restricted to the one call the reference makes:
 * XXH64(&u64, 8, seed)  (double_bloomfilter.hpp:276-281). */
uint64_t ko_xxh64_u64(uint64_t value, uint64_t seed);

/* double_bloomfilter.hpp:432-451 : the fixed seed list; returns seeds[i] (i < 110) */
uint64_t ko_bloom_seed(uint32_t i);

/* main.cpp:401-418 : Bloom sizing.  m = bits per filter (power of two), nh_ceil = hash functions
 * used in pass 1, nh_floor = hash functions tested in pass 2 (main.cpp:472 passes the double). */
void ko_bloom_params(uint64_t expected_unique, double fpr, uint64_t* m, uint32_t* nh_ceil,
                     uint32_t* nh_floor);

/* ---- rolling hash: hash_functions.cpp:4-232 (RollingHasherDual) ------------------------------ */
typedef struct {
    uint64_t q, d, di, h, m;      /* modulus, base(5), base^-1 mod q, d^(m-1) mod q, window length */
    int tbm;                      /* modulus is a power of two (hash_functions.cpp:104,129)       */
    uint64_t hf, hb, hashed;      /* forward hash, hash of the reverse complement, chars hashed   */
} ko_roller;
void ko_roller_init(ko_roller* r, uint64_t q, uint64_t k, int tbm);
void ko_roller_reset(ko_roller* r);                              /* hash_functions.cpp:235-240 */
void ko_roller_update(ko_roller* r, uint64_t in, uint64_t out);  /* hash_functions.cpp:194-208 */

/* ---- counting -------------------------------------------------------------------------------- */
/* Result of a count: n distinct canonical k-mers, sorted ascending by key.  A key is W=ceil(k/32)
 * 64-bit words, the 2k-bit big-endian integer right-aligned (word 0 holds the leading 2k mod 64
 * bits) -- the same layout as KMerFactoryCanonical2BC::blocks (kmer_factory.cpp:31-33,172-239). */
typedef struct {
    uint64_t n;             /* distinct canonical k-mers                                       */
    uint32_t W;             /* words per key                                                   */
    uint32_t k;
    uint64_t* keys;         /* n * W words                                                     */
    uint64_t* counts;       /* n true multiplicities                                           */
    uint64_t total_windows; /* input k-mers = number of complete windows seen                  */
    uint64_t invalid_bytes; /* bytes char2int mapped to 4 outside headers (excl. FASTA newline) */
} ko_counts;

/* parallel_parser.hpp:597-702 (FASTA) / :391-455 (PLAIN) scan + canonical selection, counted into
 * a private open-addressing map.  starts_in_header mirrors text_chunk::broken_header
 * (text_reader.h:22, parallel_parser.hpp:596). Returns 0, or -1 on allocation failure / bad k. */
int ko_count(const uint8_t* buf, size_t n, uint32_t k, int input_mode, int starts_in_header,
             ko_counts* out);
void ko_counts_free(ko_counts* c);

/* Count the reference would REPORT for a true multiplicity:
 * plain table uint16 wrap (parallel_parser.hpp:720-734), Kaarme 14-bit saturation (kmer.cpp:699-714) */
uint64_t ko_reported_count(uint64_t true_count, int table_mode);

/* kmer_hash_table.cpp:2022-2043 writer format: "<k chars ACGT> <count>\n" for every k-mer whose
 * reported count >= min_abundance (min_abundance 0 writes nothing, parallel_parser.hpp:860-861).
 * Lines come out sorted by k-mer (== `sort` of the reference output under LC_ALL=C).
 * Returns bytes written into dst (or needed, if dst is NULL / cap too small). */
size_t ko_format(const ko_counts* c, uint64_t min_abundance, int table_mode, char* dst, size_t cap);

/* key words -> k chars (no terminator) */
void ko_key_to_string(const uint64_t* key, uint32_t k, char* dst);

/* ---- double Bloom filter, sequential spec ---------------------------------------- */
/* main.cpp:395-461 + double_bloomfilter.hpp:276-413 + mybitarray.hpp:30-162, executed by ONE worker
 * in input order (the reference with a single worker thread is deterministic and equals this).
 * pass 1 over the whole input; returns new_in_first/new_in_second and, in f2 (caller-allocated,
 * m/8 bytes), the squeezed second filter (bit h of the filter = byte h>>3, mask 0x80>>(h&7)). */
typedef struct {
    uint64_t m;
    uint32_t nh_ceil, nh_floor;
    uint64_t new_in_first, new_in_second;
    uint64_t table_slots; /* next_prime3mod4(2*new_in_second), main.cpp:454 + parallel_parser.hpp:1590 */
} ko_bloom_stats;
int ko_bloom_pass1(const uint8_t* buf, size_t n, uint32_t k, int input_mode, uint64_t expected_unique,
                   double fpr, uint8_t* f2, ko_bloom_stats* st);
/* pass 2 admission rule (parallel_parser.hpp:2021-2026): all of the first nh_floor bits set in f2 */
int ko_bloom_admits(const uint8_t* f2, const ko_bloom_stats* st, uint64_t root);
/* full two-pass count: only admitted windows are counted */
int ko_count_bloom(const uint8_t* buf, size_t n, uint32_t k, int input_mode, uint64_t expected_unique,
                   double fpr, ko_counts* out, ko_bloom_stats* st);

/* ---- Kaarme 8-byte slot decode ---------------------------------------------------- */
/* kmer.hpp:107-123 bit layout; reconstruct_kmer_in_slot kmer_hash_table.cpp:3848-4058.
 * table: n_slots u64 slots; roots: root r occupies W words at roots[r*W].
 * Decodes the canonical string stored at `slot` into out_codes[k] (values 0..3).
 * Returns number of chain hops, or -1 on a malformed chain (unoccupied node / > k hops). */
int64_t ko_kaarme_decode(const uint64_t* table, uint64_t n_slots, const uint64_t* roots, uint32_t k,
                         uint64_t slot, uint8_t* out_codes);

#ifdef __cplusplus
}
#endif
#endif
