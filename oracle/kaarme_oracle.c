/* kaarme_oracle.c -- TEST INFRASTRUCTURE ONLY (see kaarme_oracle.h).
 *
 * CPU restatement of the reference's counting hot path, single-threaded, written from the
 * behaviour of the files cited at each function (paths relative to /root/reference).
 * Parity status: PINNED by tests/test_oracle_golden.py (golden vectors minted by the reference
 * binary + live comparison against oracle/_ref/kaarme when it is present).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs may use it.
 */
#include "kaarme_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * scalar helpers
 * ---------------------------------------------------------------------------------------------- */

/* functions_strings.cpp:56-70 */
uint32_t ko_char2int(uint8_t c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

/* functions_math.cpp:53-96: trial division by every integer 3..floor(sqrt(n)); the candidate walks
 * over odd numbers and is accepted when it is prime AND == 3 (mod 4). */
uint64_t ko_next_prime3mod4(uint64_t at_least) {
    uint64_t c = at_least;
    if (c <= 2) return 2;
    if ((c & 1) == 0) c += 1;
    for (;;) {
        uint64_t lim = (uint64_t)floor(sqrt((double)c));
        int prime = 1;
        for (uint64_t d = 3; d <= lim; d++) {
            if (c % d == 0) { prime = 0; break; }
        }
        if (prime && (c % 4 == 3)) return c;
        c += 2;
    }
}

/* functions_math.cpp:99-130 */
uint64_t ko_modinv(int64_t A, int64_t M) {
    int64_t m0 = M, y = 0, x = 1;
    if (M == 1) return 0;
    while (A > 1) {
        int64_t q = A / M, t = M;
        M = A % M; A = t;
        t = y; y = x - q * y; x = t;
    }
    if (x < 0) x += m0;
    return (uint64_t)x;
}

/* external/xxHash/xxhash.h: primes :3353-3357, round :3359-3383, avalanche :3385-3393,
 * finalize 8-byte step :3418-3424, short-input init :3485-3488. */
#define P64_1 0x9E3779B185EBCA87ULL
#define P64_2 0xC2B2AE3D27D4EB4FULL
#define P64_3 0x165667B19E3779F9ULL
#define P64_4 0x85EBCA77C2B2AE63ULL
#define P64_5 0x27D4EB2F165667C5ULL
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

uint64_t ko_xxh64_u64(uint64_t value, uint64_t seed) {
    uint64_t h = seed + P64_5 + 8;                     /* len < 32: h64 = seed + PRIME5; h64 += len */
    uint64_t k1 = rotl64(value * P64_2, 31) * P64_1;   /* XXH64_round(0, value)                    */
    h ^= k1;
    h = rotl64(h, 27) * P64_1 + P64_4;
    h ^= h >> 33; h *= P64_2;                          /* avalanche                                */
    h ^= h >> 29; h *= P64_3;
    h ^= h >> 32;
    return h;
}

/* double_bloomfilter.hpp:434-444 */
static const uint64_t KO_SEEDS[110] = {
    2411, 3253, 1061, 1129, 2269, 7309, 3491, 8237, 6359, 8779, 6553, 5443, 2447, 8999, 8623, 5779,
    1879, 2357, 5087, 5393, 2203, 8597, 8629, 7727, 2819, 1789, 7757, 6079, 9371, 2957, 2389, 4133,
    4931, 2083, 8291, 1151, 4759, 7649, 6803, 1753, 9613, 1979, 1877, 5479, 4799, 5303, 1759, 4451,
    7841, 3461, 2207, 1289, 5233, 6823, 7043, 3251, 8039, 4519, 2551, 5693, 7681, 1607, 4679, 4729,
    8231, 4139, 7457, 6221, 2377, 7151, 3083, 6947, 7331, 3947, 6011, 7753, 2843, 3191, 7993, 4943,
    5801, 9901, 4001, 1933, 7523, 5273, 1721, 1093, 2579, 9719, 4481, 2417, 9341, 9137, 5113, 9719,
    5399, 5231, 1979, 6701, 4133, 1723, 1931, 3257, 8861, 8539, 4877, 2207, 7151, 5279};
uint64_t ko_bloom_seed(uint32_t i) { return KO_SEEDS[i % 110]; }

/* main.cpp:401-418 */
void ko_bloom_params(uint64_t U, double fpr, uint64_t* m, uint32_t* nh_ceil, uint32_t* nh_floor) {
    double bits_min = (-(double)U * log(fpr)) / pow(log(2.0), 2.0);
    double h = (bits_min / (double)U) * log(2.0);
    uint64_t p2 = 2;
    while (p2 < (uint64_t)bits_min) p2 *= 2;
    *m = p2;
    *nh_ceil = (uint32_t)ceil(h);
    *nh_floor = (uint32_t)h; /* main.cpp:472 double -> uint64_t parameter conversion truncates */
}

/* ------------------------------------------------------------------------------------------------
 * rolling hash (hash_functions.cpp)
 * ---------------------------------------------------------------------------------------------- */
typedef unsigned __int128 u128;

/* ctor: hash_functions.cpp:4-26 (q,m) and :52-75 (explicit inverse). h = d^(m-1) mod q (:24-25). */
void ko_roller_init(ko_roller* r, uint64_t q, uint64_t k, int tbm) {
    r->q = q; r->d = 5; r->m = k; r->tbm = tbm;
    r->di = ko_modinv(5, (int64_t)q);
    r->h = 1;
    for (uint64_t i = 0; i + 1 < k; i++) r->h = (uint64_t)(((u128)r->h * r->d) % q);
    ko_roller_reset(r);
}
void ko_roller_reset(ko_roller* r) { r->hf = 0; r->hb = 0; r->hashed = 0; }

/* hash_functions.cpp:194-208 dispatch; :102-124 warm-up (in only); :127-192 steady state (in+out).
 * `& (q-1)` and `% q` coincide for power-of-two q, so one code path serves tbm and !tbm. */
void ko_roller_update(ko_roller* r, uint64_t in, uint64_t out) {
    const uint64_t q = r->q;
    if (r->hashed < r->m) {
        r->hf = (uint64_t)(((u128)r->d * r->hf + in) % q);
        u128 radd = 3 - in;                                   /* reverse_int(in)                  */
        for (uint64_t i = 0; i < r->hashed; i++) radd = (radd * r->d) % q;
        r->hb = (uint64_t)(((u128)r->hb + radd) % q);
    } else {
        u128 baseline = ((u128)r->d * r->hf + in) % q;
        u128 rem = ((u128)r->d * r->h * out) % q;
        r->hf = (uint64_t)(rem > baseline ? (baseline + q) - rem : baseline - rem);
        baseline = r->hb;
        rem = 3 - out;
        baseline = rem > baseline ? (baseline + q) - rem : baseline - rem;
        baseline = (baseline * (u128)r->di) % q;
        r->hb = (uint64_t)((((u128)(3 - in)) * r->h + baseline) % q);
    }
    r->hashed = r->hashed + 1 < r->m ? r->hashed + 1 : r->m;
}

/* ------------------------------------------------------------------------------------------------
 * k-mer window (forward + reverse complement), W-word keys
 * ---------------------------------------------------------------------------------------------- */
#define KO_MAXW 16 /* k <= 512 */

typedef struct {
    uint32_t k, W;
    uint64_t topmask;       /* mask of the used bits of word 0                                   */
    uint32_t topshift;      /* bit position (in word 0) of the leading character                 */
    uint64_t f[KO_MAXW];    /* forward window, right-aligned                                      */
    uint64_t r[KO_MAXW];    /* reverse complement of the window                                  */
    uint32_t chars;         /* characters currently in the window (<= k)                         */
} ko_window;

static void win_init(ko_window* w, uint32_t k) {
    memset(w, 0, sizeof(*w));
    w->k = k; w->W = (k + 31) / 32;
    uint32_t top_chars = k - 32 * (w->W - 1);   /* 1..32 characters live in word 0 */
    w->topmask = top_chars == 32 ? ~0ULL : ((1ULL << (2 * top_chars)) - 1);
    w->topshift = 2 * (top_chars - 1);
}
static void win_reset(ko_window* w) {
    memset(w->f, 0, sizeof(w->f)); memset(w->r, 0, sizeof(w->r)); w->chars = 0;
}
/* forward: shift left by one character, append c (parallel_parser.hpp:646-654 on bytes,
 * kmer_factory.cpp:172-205 on 64-bit blocks).  reverse complement: shift right by one character and
 * put 3-c in front (kmer_factory.cpp:207-217; parallel_parser.hpp:660-681 rebuilds it from scratch
 * per base -- same value). Returns the character pushed off the forward window. */
static uint32_t win_push(ko_window* w, uint32_t c) {
    const uint32_t W = w->W;
    uint32_t out = (uint32_t)((w->f[0] >> w->topshift) & 3);
    for (uint32_t i = 0; i + 1 < W; i++) w->f[i] = (w->f[i] << 2) | (w->f[i + 1] >> 62);
    w->f[W - 1] = (w->f[W - 1] << 2) | c;
    w->f[0] &= w->topmask;
    for (uint32_t i = W - 1; i > 0; i--) w->r[i] = (w->r[i] >> 2) | (w->r[i - 1] << 62);
    w->r[0] = (w->r[0] >> 2) | ((uint64_t)(3 - c) << w->topshift);
    if (w->chars < w->k) w->chars++;
    return out;
}
/* parallel_parser.hpp:686-702 / kmer_factory.cpp:219-233: lexicographic min, forward wins ties */
static int win_forward_is_canonical(const ko_window* w) {
    for (uint32_t i = 0; i < w->W; i++) {
        if (w->r[i] < w->f[i]) return 0;
        if (w->r[i] > w->f[i]) return 1;
    }
    return 1;
}

/* ------------------------------------------------------------------------------------------------
 * private open-addressing map (oracle bookkeeping, NOT a restatement of the reference table: the
 * parity contract is the multiset of canonical k-mers)
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    uint32_t W;
    uint64_t cap, n; /* cap is a power of two */
    uint64_t* keys;
    uint64_t* counts;
} ko_map;

static uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
static uint64_t key_hash(const uint64_t* key, uint32_t W) {
    uint64_t h = 0x9E3779B97F4A7C15ULL;
    for (uint32_t i = 0; i < W; i++) h = mix64(h ^ key[i]) + 0x9E3779B97F4A7C15ULL * (i + 1);
    return h;
}
static int map_init(ko_map* m, uint32_t W, uint64_t cap) {
    m->W = W; m->cap = cap; m->n = 0;
    m->keys = (uint64_t*)malloc(sizeof(uint64_t) * cap * W);
    m->counts = (uint64_t*)calloc(cap, sizeof(uint64_t));
    if (!m->keys || !m->counts) { free(m->keys); free(m->counts); return -1; }
    return 0;
}
static void map_free(ko_map* m) { free(m->keys); free(m->counts); m->keys = NULL; m->counts = NULL; }
static int map_add(ko_map* m, const uint64_t* key, uint64_t inc);
static int map_grow(ko_map* m) {
    ko_map b;
    if (map_init(&b, m->W, m->cap * 2)) return -1;
    for (uint64_t s = 0; s < m->cap; s++)
        if (m->counts[s]) map_add(&b, m->keys + s * m->W, m->counts[s]);
    map_free(m);
    *m = b;
    return 0;
}
static int map_add(ko_map* m, const uint64_t* key, uint64_t inc) {
    if ((m->n + 1) * 10 > m->cap * 6 && map_grow(m)) return -1;
    const uint32_t W = m->W;
    uint64_t s = key_hash(key, W) & (m->cap - 1);
    for (;;) {
        if (m->counts[s] == 0) {
            memcpy(m->keys + s * W, key, sizeof(uint64_t) * W);
            m->counts[s] = inc; m->n++;
            return 0;
        }
        if (memcmp(m->keys + s * W, key, sizeof(uint64_t) * W) == 0) { m->counts[s] += inc; return 0; }
        s = (s + 1) & (m->cap - 1);
    }
}

/* sort helper: order slot indices by key, ascending (word 0 most significant) */
static _Thread_local uint32_t g_sortW;              /* comparator context: per thread, so concurrent ko_count calls are safe */
static _Thread_local const uint64_t* g_sortkeys;
static int cmp_slot(const void* a, const void* b) {
    const uint64_t* ka = g_sortkeys + (*(const uint64_t*)a) * g_sortW;
    const uint64_t* kb = g_sortkeys + (*(const uint64_t*)b) * g_sortW;
    for (uint32_t i = 0; i < g_sortW; i++) {
        if (ka[i] < kb[i]) return -1;
        if (ka[i] > kb[i]) return 1;
    }
    return 0;
}
static int map_to_sorted(ko_map* m, ko_counts* out) {
    uint64_t n = m->n, j = 0;
    uint64_t* idx = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
    out->keys = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1) * m->W);
    out->counts = (uint64_t*)malloc(sizeof(uint64_t) * (n ? n : 1));
    if (!idx || !out->keys || !out->counts) { free(idx); return -1; }
    for (uint64_t s = 0; s < m->cap; s++) if (m->counts[s]) idx[j++] = s;
    g_sortW = m->W; g_sortkeys = m->keys;
    qsort(idx, n, sizeof(uint64_t), cmp_slot);
    for (uint64_t i = 0; i < n; i++) {
        memcpy(out->keys + i * m->W, m->keys + idx[i] * m->W, sizeof(uint64_t) * m->W);
        out->counts[i] = m->counts[idx[i]];
    }
    out->n = n; out->W = m->W;
    free(idx);
    return 0;
}

/* ------------------------------------------------------------------------------------------------
 * the scanner: one callback per complete window
 * ---------------------------------------------------------------------------------------------- */
typedef int (*ko_window_fn)(void* user, const ko_window* w, int forward_is_canonical,
                            const ko_roller* roller);

/* FASTA branch parallel_parser.hpp:597-638 (same scanner in every functor: :1398-1439, :1958-1999,
 * :2488-2533, :2853-2887); PLAIN branch :391-400.  The optional roller is updated exactly as the
 * functor does (update_rolling_hash(new, dropped) per valid base, reset on breaks). */
static int ko_scan(const uint8_t* buf, size_t n, uint32_t k, int input_mode, int starts_in_header,
                   ko_roller* roller, ko_window_fn fn, void* user, uint64_t* total_windows,
                   uint64_t* invalid_bytes) {
    ko_window w;
    if (k == 0 || (k + 31) / 32 > KO_MAXW) return -1;
    win_init(&w, k);
    size_t i = 0;
    int parsing_header = starts_in_header;
    uint64_t tw = 0, inv = 0;
    while (i < n) {
        if (input_mode == KO_FASTA) {
            if (buf[i] == '>') parsing_header = 1;
            if (parsing_header) {
                while (i < n && buf[i] != '\n') i++;
                i++;
                parsing_header = 0;
                win_reset(&w);
                if (roller) ko_roller_reset(roller);
                continue;
            }
            if (buf[i] == '\n') { i++; continue; }
        }
        uint32_t c = ko_char2int(buf[i]);
        if (c > 3) {
            if (buf[i] != '\n') inv++;
            win_reset(&w);
            if (roller) ko_roller_reset(roller);
        } else {
            uint32_t out = win_push(&w, c);
            if (roller) ko_roller_update(roller, c, out);
            if (w.chars >= k) {
                tw++;
                int rc = fn(user, &w, win_forward_is_canonical(&w), roller);
                if (rc) return rc;
            }
        }
        i++;
    }
    if (total_windows) *total_windows = tw;
    if (invalid_bytes) *invalid_bytes = inv;
    return 0;
}

static int cb_count(void* user, const ko_window* w, int fwd, const ko_roller* roller) {
    (void)roller;
    return map_add((ko_map*)user, fwd ? w->f : w->r, 1);
}

int ko_count(const uint8_t* buf, size_t n, uint32_t k, int input_mode, int starts_in_header,
             ko_counts* out) {
    ko_map m;
    memset(out, 0, sizeof(*out));
    if (k == 0 || (k + 31) / 32 > KO_MAXW) return -1;
    if (map_init(&m, (k + 31) / 32, 1 << 16)) return -1;
    int rc = ko_scan(buf, n, k, input_mode, starts_in_header, NULL, cb_count, &m, &out->total_windows,
                     &out->invalid_bytes);
    if (rc == 0) rc = map_to_sorted(&m, out);
    out->k = k;
    map_free(&m);
    return rc;
}

void ko_counts_free(ko_counts* c) {
    free(c->keys); free(c->counts);
    c->keys = NULL; c->counts = NULL; c->n = 0;
}

/* SURVEY A.1.5. plain: `counts[slot] += 1` on a uint16_t (parallel_parser.hpp:732), and a slot whose
 * counter wrapped to 0 is re-initialised to 1 by the next hit (:722-727) => N mod 65536.
 * Kaarme: increase_count() stops at 16383 (kmer.cpp:699-714). */
uint64_t ko_reported_count(uint64_t true_count, int table_mode) {
    if (table_mode == KO_TABLE_PLAIN) return true_count & 0xFFFFULL;
    if (table_mode == KO_TABLE_KAARME) return true_count > 16383 ? 16383 : true_count;
    return true_count;
}

void ko_key_to_string(const uint64_t* key, uint32_t k, char* dst) {
    const uint32_t W = (k + 31) / 32;
    for (uint32_t j = 0; j < k; j++) {
        uint32_t pos = k - 1 - j;            /* character index from the right end */
        uint64_t word = key[W - 1 - pos / 32];
        dst[j] = "ACGT"[(word >> (2 * (pos % 32))) & 3];
    }
}

/* kmer_hash_table.cpp:2013-2050 (and the Kaarme writer :4318-4524 emits the same line format) */
size_t ko_format(const ko_counts* c, uint64_t min_abundance, int table_mode, char* dst, size_t cap) {
    size_t need = 0;
    if (min_abundance == 0) return 0; /* parallel_parser.hpp:860-861 */
    char num[32];
    for (uint64_t i = 0; i < c->n; i++) {
        uint64_t rep = ko_reported_count(c->counts[i], table_mode);
        if (rep < min_abundance) continue;
        int nd = 0; uint64_t v = rep;
        do { num[nd++] = (char)('0' + v % 10); v /= 10; } while (v);
        size_t line = (size_t)c->k + 1 + (size_t)nd + 1;
        if (dst && need + line <= cap) {
            char* p = dst + need;
            ko_key_to_string(c->keys + i * c->W, c->k, p);
            p += c->k; *p++ = ' ';
            while (nd) *p++ = num[--nd];
            *p = '\n';
        }
        need += line;
    }
    return need;
}

/* ------------------------------------------------------------------------------------------------
 * double Bloom filter -- sequential spec
 * ---------------------------------------------------------------------------------------------- */
/* MyAtomicBitArrayFT (mybitarray.hpp:30-125): 2m bits, MSB-first within a byte, physically split in
 * two halves -- a split that is invisible to a sequential reader, so one flat array serves. */
static inline int bit_test(const uint8_t* a, uint64_t i) { return (a[i >> 3] >> (7 - (i & 7))) & 1; }
static inline void bit_set(uint8_t* a, uint64_t i) { a[i >> 3] |= (uint8_t)(0x80u >> (i & 7)); }

typedef struct {
    uint8_t* bits; /* interleaved: bit 2h = filter 1, bit 2h+1 = filter 2 (double_bloomfilter.hpp:303-368) */
    uint64_t mask;
    uint32_t nh;
    uint64_t new_in_first, new_in_second;
} ko_dbf;

/* insert_in_first / insert_in_second (double_bloomfilter.hpp:339-368): set every bit that tests 0 and
 * report whether THIS call flipped exactly (nh - already_set) bits.  Executed by a single thread a
 * set() never loses a race, but the report is still false when two of the nh hash values coincide
 * (the second one finds the bit already set), which sends the k-mer down the "race" branch. */
static int dbf_set_all(ko_dbf* f, const uint64_t* hv, uint32_t odd, uint32_t already) {
    uint32_t mine = 0;
    for (uint32_t i = 0; i < f->nh; i++) {
        uint64_t b = 2 * hv[i] + odd;
        if (!bit_test(f->bits, b)) { bit_set(f->bits, b); mine++; }
    }
    return mine == f->nh - already;
}

/* double_bloomfilter.hpp:371-413 (insertion_process) executed by one thread in input order */
static void dbf_insert(ko_dbf* f, uint64_t root) {
    uint64_t hv[110];
    uint32_t in2 = 0, in1 = 0;
    for (uint32_t i = 0; i < f->nh; i++) hv[i] = ko_xxh64_u64(root, KO_SEEDS[i]) & f->mask;
    for (uint32_t i = 0; i < f->nh; i++) in2 += bit_test(f->bits, 2 * hv[i] + 1);
    if (in2 == f->nh) return;
    for (uint32_t i = 0; i < f->nh; i++) in1 += bit_test(f->bits, 2 * hv[i]);
    if (in1 == f->nh) {
        if (dbf_set_all(f, hv, 1, in2)) f->new_in_second++;
    } else if (dbf_set_all(f, hv, 0, in1)) {
        f->new_in_first++;
    } else if (dbf_set_all(f, hv, 1, in2)) { /* :401-411 "inserted by someone else" branch */
        f->new_in_second++;
    }
}

/* parallel_parser.hpp:2889-2894: root = min(Hb, Hf) of the mod-2^54 base-5 rolling hash */
static int cb_bloom1(void* user, const ko_window* w, int fwd, const ko_roller* r) {
    (void)w; (void)fwd;
    dbf_insert((ko_dbf*)user, r->hf < r->hb ? r->hf : r->hb);
    return 0;
}

int ko_bloom_pass1(const uint8_t* buf, size_t n, uint32_t k, int input_mode, uint64_t expected_unique,
                   double fpr, uint8_t* f2, ko_bloom_stats* st) {
    ko_dbf f;
    ko_roller roller;
    memset(st, 0, sizeof(*st));
    ko_bloom_params(expected_unique, fpr, &st->m, &st->nh_ceil, &st->nh_floor);
    f.mask = st->m - 1; f.nh = st->nh_ceil; f.new_in_first = f.new_in_second = 0;
    f.bits = (uint8_t*)calloc((size_t)(st->m / 4) + 1, 1); /* 2m bits */
    if (!f.bits) return -1;
    ko_roller_init(&roller, 1ULL << 54, k, 0); /* main.cpp:433-435; 5-arg ctor => tbm=false, parallel_parser.hpp:2799 */
    int rc = ko_scan(buf, n, k, input_mode, 0, &roller, cb_bloom1, &f, NULL, NULL);
    st->new_in_first = f.new_in_first;
    st->new_in_second = f.new_in_second;
    st->table_slots = ko_next_prime3mod4(2 * f.new_in_second);
    /* squeeze (mybitarray.hpp:127-162): keep the odd bits */
    if (f2) {
        memset(f2, 0, (size_t)(st->m / 8) + 1);
        for (uint64_t h = 0; h < st->m; h++) if (bit_test(f.bits, 2 * h + 1)) bit_set(f2, h);
    }
    free(f.bits);
    return rc;
}

int ko_bloom_admits(const uint8_t* f2, const ko_bloom_stats* st, uint64_t root) {
    for (uint32_t i = 0; i < st->nh_floor; i++)
        if (!bit_test(f2, ko_xxh64_u64(root, KO_SEEDS[i]) & (st->m - 1))) return 0;
    return 1;
}

typedef struct { ko_map* m; const uint8_t* f2; const ko_bloom_stats* st; } ko_p2;
static int cb_bloom2(void* user, const ko_window* w, int fwd, const ko_roller* r) {
    ko_p2* p = (ko_p2*)user;
    if (!ko_bloom_admits(p->f2, p->st, r->hf < r->hb ? r->hf : r->hb)) return 0;
    return map_add(p->m, fwd ? w->f : w->r, 1);
}

int ko_count_bloom(const uint8_t* buf, size_t n, uint32_t k, int input_mode, uint64_t expected_unique,
                   double fpr, ko_counts* out, ko_bloom_stats* st) {
    ko_map m;
    ko_roller roller;
    memset(out, 0, sizeof(*out));
    uint64_t mm; uint32_t a, b;
    ko_bloom_params(expected_unique, fpr, &mm, &a, &b);
    uint8_t* f2 = (uint8_t*)malloc((size_t)(mm / 8) + 1);
    if (!f2) return -1;
    int rc = ko_bloom_pass1(buf, n, k, input_mode, expected_unique, fpr, f2, st);
    if (rc == 0 && map_init(&m, (k + 31) / 32, 1 << 16) == 0) {
        ko_p2 p = {&m, f2, st};
        ko_roller_init(&roller, 1ULL << 54, k, 1); /* parallel_parser.hpp:1701 */
        rc = ko_scan(buf, n, k, input_mode, 0, &roller, cb_bloom2, &p, &out->total_windows,
                     &out->invalid_bytes);
        if (rc == 0) rc = map_to_sorted(&m, out);
        out->k = k;
        map_free(&m);
    } else if (rc == 0) rc = -1;
    free(f2);
    return rc;
}

/* ------------------------------------------------------------------------------------------------
 * Kaarme slot decode (kmer.hpp:107-123, kmer_hash_table.cpp:3848-4058, :3073-3093)
 * ---------------------------------------------------------------------------------------------- */
#define KS_OCC(d) ((d) & 1ULL)
#define KS_HASPRED(d) (((d) >> 1) & 1ULL)
#define KS_SELF_FWD(d) (((d) >> 4) & 1ULL)
#define KS_PRED_FWD(d) (((d) >> 5) & 1ULL)
#define KS_RIGHT(d) (((d) >> 8) & 3ULL)
#define KS_LEFT(d) (((d) >> 10) & 3ULL)
#define KS_PTR(d) ((d) >> 26)

static uint32_t root_char(const uint64_t* roots, uint64_t r, uint32_t k, uint32_t W, int pos_from_left) {
    uint32_t pos = k - 1 - (uint32_t)pos_from_left;
    uint64_t word = roots[r * W + (W - 1 - pos / 32)];
    return (uint32_t)((word >> (2 * (pos % 32))) & 3);
}

int64_t ko_kaarme_decode(const uint64_t* table, uint64_t n_slots, const uint64_t* roots, uint32_t k,
                         uint64_t slot, uint8_t* out) {
    const uint32_t W = (k + 31) / 32;
    int L = 0, R = (int)k - 1, Lc = 0, Rc = (int)k - 1, pir = 0;
    uint64_t pos = slot;
    int64_t hops = 0;
    if (slot >= n_slots || !KS_OCC(table[slot])) return -1;
    for (;;) {
        uint64_t d = table[pos];
        if (!KS_OCC(d)) return -1;
        if (!KS_HASPRED(d)) break;
        if (L == Lc) { out[L] = (uint8_t)(pir ? 3 - KS_RIGHT(d) : KS_LEFT(d)); L++; if (L > R) return hops; }
        if (R == Rc) { out[R] = (uint8_t)(pir ? 3 - KS_LEFT(d) : KS_RIGHT(d)); R--; if (L > R) return hops; }
        int s = (int)KS_SELF_FWD(d), p = (int)KS_PRED_FWD(d);
        int shift = s ? -1 : 1;         /* :3923-4001 the eight cases collapse to this */
        if (pir) shift = -shift;
        Lc += shift; Rc += shift;
        if (s != p) pir = !pir;
        pos = KS_PTR(d);
        if (pos >= n_slots) return -1;
        if (++hops > (int64_t)n_slots) return -1; /* acyclic chains are shorter than the table */
    }
    {
        uint64_t r = KS_PTR(table[pos]);
        int Ls = L - Lc;
        if (!pir) { for (int a = L, b = Ls; a <= R; a++, b++) { if (b < 0 || b >= (int)k) return -1; out[a] = (uint8_t)root_char(roots, r, k, W, b); } }
        else { for (int a = L, b = (int)k - Ls - 1; a <= R; a++, b--) { if (b < 0 || b >= (int)k) return -1; out[a] = (uint8_t)(3 - root_char(roots, r, k, W, b)); } }
    }
    return hops;
}
