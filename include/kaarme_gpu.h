/* kaarme_gpu.h -- C ABI of libkaarme_gpu.so, the B200 (sm_100a) k-mer counting path.
 *
 * The reference (Denopia/canonical-k-mer-hash-table, "Kaarme") has no plugin/FFI interface: its
 * counting path sits behind C++ functors called once from main.cpp:442-536.  This header is the
 * boundary a maintainer would bind instead of those functors; each entry point cites what it replaces
 * (paths relative to the reference root).  INTEGRATION.md shows the reference-side stub.
 *
 *   reference                                                     this ABI
 *   ------------------------------------------------------------  -----------------------------------
 *   struct arguments                         main.cpp:70-96        kg_config
 *   new BasicAtomicFlagHashTableLong /       parallel_parser.hpp:236-239,
 *       PointerHashTableCanonicalAV          :1189-1196            kg_create + kg_pass_begin(COUNT)
 *   new DoubleAtomicDoubleBloomFilter        main.cpp:395-431      kg_create (use_bloom)
 *   text_chunk + read_chunk_from_file        text_reader.h:17-226  kg_stream_begin + kg_feed
 *   hash_kmers lambda (scan, window,         parallel_parser.hpp:344-809 (-m 0), :1300-1500 (-m 2),
 *       canonical, hash, insert)             :1700-2242, :2380-2672 (Bloom pass 2)   kg_feed in KG_PASS_COUNT
 *   ..._BLOOM_FILTERING functor              parallel_parser.hpp:2678-2974           kg_feed in KG_PASS_BLOOM
 *   get_new_in_second / resize / 2x sizing   main.cpp:454-461      kg_pass_end(BLOOM) -> kg_pass_begin(COUNT)
 *   write_kmers / write_kmers_on_disk_...    kmer_hash_table.cpp:2013-2050, :4318-4524   kg_export
 *   Kaarme 8-byte slots + secondary array    kmer.hpp:103-149, kmer_hash_table.cpp:2128-2282   kg_compact
 *
 * Conventions: plain pointers and sizes only; every function returns a kg_status (0 = ok) and never
 * aborts; the library owns all device memory; one host thread drives a context (internal CUDA streams
 * give the concurrency); host batches belong to the caller again as soon as kg_feed returns.
 * There is NO CPU fallback: without a usable sm_100 device kg_create fails with KG_ECUDA.
 */
#ifndef KAARME_GPU_H
#define KAARME_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KG_ABI_VERSION 2

typedef enum kg_status {
    KG_OK = 0,
    KG_EBADARG = 1,     /* bad configuration / call order                                           */
    KG_ECUDA = 2,       /* CUDA runtime error or no usable device (kg_last_error has the text)      */
    KG_ETABLE_FULL = 3, /* open-addressing table full (reference: parallel_parser.hpp:742-746 prints
                           and silently truncates; kmer_hash_table.cpp:2552-2556 exit(1))            */
    KG_ENCCL = 4,       /* NCCL error                                                               */
    KG_ENOMEM = 5,      /* device or pinned-host allocation failed                                  */
    KG_ESINK = 6        /* export sink returned non-zero                                            */
} kg_status;

/* main.cpp:179-188 input_mode */
#define KG_INPUT_FASTA 0
#define KG_INPUT_PLAIN 2
/* main.cpp:137 hash_table_mode */
#define KG_TABLE_PLAIN 0
#define KG_TABLE_KAARME 2
/* passes over the input (main.cpp:442 Bloom pass, :468-536 count pass) */
#define KG_PASS_BLOOM 1
#define KG_PASS_COUNT 2
/* kg_feed flags */
#define KG_FEED_CONTEXT 1u /* bytes only warm the k-1 window / header state; no k-mer ending inside
                              them is counted (the overlap of text_reader.h:141-151,210)            */
/* count_mode for kg_export: which counter width the reported counts emulate */
#define KG_COUNT_EXACT 0    /* 32-bit true multiplicity                                             */
#define KG_COUNT_REFERENCE 1 /* -m 0: N mod 65536 (uint16 wrap); -m 2: min(N, 16383)                */

#define KG_MAX_K 256 /* keys are ceil(k/32) 64-bit words, up to 8 */

typedef struct kg_ctx kg_ctx;

typedef struct kg_config {
    uint32_t abi_version;     /* KG_ABI_VERSION                                                     */
    uint32_t k;               /* KLEN                      main.cpp:135                            */
    int32_t table_mode;       /* -m   KG_TABLE_*           main.cpp:137                            */
    int32_t input_mode;       /* KG_INPUT_*                main.cpp:27-68,179-188                  */
    uint64_t min_slots;       /* -s   (ignored with use_bloom)   main.cpp:145                      */
    int32_t use_bloom;        /* -b                        main.cpp:141                            */
    int32_t device;           /* CUDA device ordinal                                                */
    double fpr;               /* -f                        main.cpp:142                            */
    uint64_t expected_unique; /* -u                        main.cpp:146                            */
    uint64_t batch_bytes;     /* raw bytes per device batch; 0 = default (128 MiB)                  */
    int32_t rank;             /* hash-sharded multi-GPU: this context's shard                       */
    int32_t world;            /* number of shards (1 = single GPU)                                  */
    uint32_t partitions;      /* L2-blocked insert: the windows of every batch are bucketed by the hash of their
                                 minimizer into this many contiguous regions of the shard's table / filter before
                                 inserting.  0 = choose per pass (~24 MiB regions; direct insert when the structure
                                 is small), 1 = always insert directly (one GPU), > 1 = as given
                                 (world * partitions <= 1024)                                                 */
    uint32_t reserved;        /* flags: KG_CFG_* below (0 = the defaults)                                              */
} kg_config;

/* kg_config.reserved flags */
#define KG_CFG_REFERENCE_BLOOM 1u /* -b reproduces the reference's double
                                     Bloom filter bit for bit as ONE worker thread builds it -- same hash functions
                                     (base-5 rolling hash mod 2^54, XXH64 with the reference's seeds), same
                                     new_in_first / new_in_second, same table size, same false positives at -a 1 --
                                     instead of the blocked filter with its own hash.  One GPU, one stream per pass,
                                     m <= 2^31 bits, < 2^32 bases.  SURVEY.md section 8f-4; DESIGN.md section 9.        */

typedef struct kg_pass_stats {
    uint64_t input_kmers;    /* complete windows seen by THIS context in the pass ("input k-mers")  */
    uint64_t inserted_kmers; /* k-mer occurrences inserted into THIS shard (after exchange / Bloom) */
    uint64_t distinct;       /* occupied table slots (count pass)                                   */
    uint64_t table_slots;    /* next_prime3mod4(...)   parallel_parser.hpp:236, main.cpp:454        */
    uint64_t new_in_first;   /* Bloom pass: double_bloomfilter.hpp:389-396                          */
    uint64_t new_in_second;
    uint64_t bloom_bits;     /* m, bits per filter (main.cpp:404-410)                               */
    uint32_t bloom_hashes;   /* ceil(h)  (main.cpp:417)                                            */
    uint32_t partitions;     /* local table/filter regions the pass bucketed its batches into (1 = direct)  */
    uint64_t raw_bytes;      /* bytes fed in the pass                                               */
    uint64_t bases;          /* valid bases packed                                                  */
    double device_ms;        /* CUDA-event time, first kernel of the pass to the last               */
    double parse_ms, count_ms, exchange_ms; /* per-stage CUDA-event sums (0 when not measured)      */
    double insert_ms;        /* CUDA-event time of the table/filter-updating kernel alone (kg_count_kernel on
                                the direct path, kg_skm_insert on the bucketed path), summed over
                                its launches; count_ms additionally holds the bucketing kernel          */
    uint64_t insert_launches;/* launches of that kernel in the pass                                         */
} kg_pass_stats;

typedef struct kg_compact_stats {
    uint64_t kmers;          /* k-mers stored                                                       */
    uint64_t roots;          /* k-mers stored in full (secondary array), kmer_hash_table.cpp:2239   */
    uint64_t bytes;          /* 8*kmers + 8*W*roots                                                 */
    uint64_t reference_bytes;/* 8*table_slots + (8*W+1)*roots, what the reference would hold        */
    uint64_t max_chain;      /* longest predecessor chain                                           */
    double device_ms;
} kg_compact_stats;

/* Export sink: n records; keys = n*W words (W = ceil(k/32); 2k-bit big-endian integer, right-aligned,
 * word 0 most significant -- KMerFactoryCanonical2BC layout, kmer_factory.cpp:31-33), counts = n.
 * Return non-zero to abort (kg_export then returns KG_ESINK). Buffers are valid only during the call. */
typedef int (*kg_sink_fn)(void* user, const uint64_t* keys, const uint32_t* counts, size_t n);

/* Text sink: `bytes` bytes holding `records` complete output lines "<k chars ACGT> <count>\n" -- the writer format
 * of kmer_hash_table.cpp:2022-2043 -- formatted on the GPU; the host only has to write(2) them.  Line order is
 * unspecified (as in the reference).  Return non-zero to abort.  The buffer is valid only during the call. */
typedef int (*kg_text_sink_fn)(void* user, const char* text, size_t bytes, size_t records);

/* ---- lifecycle -------------------------------------------------------------------------------- */
int kg_abi_version(void);
const char* kg_strerror(int status);
const char* kg_last_error(const kg_ctx* ctx); /* text of the last CUDA/NCCL failure on this context */

/* Number of usable (compute capability 10.x) devices; 0 => the product path cannot run. */
int kg_device_count(int* count);

/* replaces the table / filter constructors. Allocates streams, staging and (use_bloom) the filter. */
int kg_create(const kg_config* cfg, kg_ctx** out);
int kg_destroy(kg_ctx* ctx);

/* pinned host memory for batches (the text_chunk buffers of parallel_parser.hpp:294-296) */
int kg_host_alloc(size_t bytes, void** out);
int kg_host_free(void* p);

/* ---- multi-GPU (one context per GPU; every shard owns a disjoint set of minimizer buckets) ----- */
/* The reference is one process on one table (main.cpp:442-536); sharding has no counterpart there.  Canonical k-mers
 * are owned by the hash of their minimizer; every GPU parses its own slice of the input and turns it into 8-byte run
 * descriptors per owner.  kg_comm_init (collective) creates the NCCL communicator -- it carries one word per round --
 * and maps every rank's two batch slots into every rank (CUDA IPC between processes, peer access between the contexts
 * of one process): the packed 2-bit reads then travel by copy-engine pulls over NVLink and the insert kernels read the
 * peers' descriptors in place.  All ranks must use the same k and batch_bytes.                                   */
#define KG_UNIQUE_ID_BYTES 256
int kg_comm_unique_id(void* id_out);                                   /* rank 0, then broadcast    */
int kg_comm_init(kg_ctx* ctx, const void* id, int rank, int world);    /* collective                */

/* ---- passes ------------------------------------------------------------------------------------ */
/* KG_PASS_BLOOM: clears the filters.  KG_PASS_COUNT: allocates and clears the table with
 * next_prime3mod4(min_slots) slots, or next_prime3mod4(2*new_in_second) after a Bloom pass.
 * Several contexts in ONE process (a host thread per GPU): let every context return from kg_pass_begin
 * before any of them feeds -- it allocates, and an allocation must not wait on a peer whose collective
 * is already in flight.                                                        */
int kg_pass_begin(kg_ctx* ctx, int pass);
/* Start an independent byte stream (a file, or one rank's slice of it). starts_in_header mirrors
 * text_chunk::broken_header (text_reader.h:22).                                                    */
int kg_stream_begin(kg_ctx* ctx, int starts_in_header);
/* Consecutive bytes of the current stream from HOST memory (pinned or pageable). Window and header
 * state carry across calls, so batch boundaries are invisible (every k-mer of every record is seen
 * exactly once -- the contract text_reader.h:91-226 implements with overlaps).                     */
int kg_feed(kg_ctx* ctx, const uint8_t* bytes, size_t n, uint32_t flags);
/* Same, bytes already resident in device memory of ctx's device.                                   */
int kg_feed_device(kg_ctx* ctx, const void* device_bytes, size_t n, uint32_t flags);
/* Drains the pipeline (multi-GPU: collective), returns the pass statistics.                        */
int kg_pass_end(kg_ctx* ctx, kg_pass_stats* stats);

/* ---- after the count pass ---------------------------------------------------------------------- */
/* -m 2: build the Kaarme representation (8-byte slot per k-mer + roots) from the counted table.  With several
 * GPUs every shard builds its own self-contained structure: a k-mer whose predecessor is owned by another shard
 * becomes a root.                                                                                   */
int kg_compact(kg_ctx* ctx, kg_compact_stats* stats);
/* Stream every k-mer whose reported count >= min_abundance to the sink (min_abundance 0 => nothing,
 * parallel_parser.hpp:860-861).  In KG_TABLE_KAARME mode after kg_compact the k-mers are DECODED from
 * the compact structure (chain walk, kmer_hash_table.cpp:3848-4058), not read from the plain table. */
int kg_export(kg_ctx* ctx, uint64_t min_abundance, int count_mode, kg_sink_fn sink, void* user);
/* Same selection, but the records reach the sink as finished text lines (replaces the char-by-char ofstream loops of
 * write_kmers, kmer_hash_table.cpp:2013-2050, and write_kmers_on_disk_separately_even_faster, :4318-4524).     */
int kg_export_text(kg_ctx* ctx, uint64_t min_abundance, int count_mode, kg_text_sink_fn sink, void* user);
/* Copy the compact structure to host memory after kg_compact: slots = kmers words (kmer.hpp:108-123 layout,
 * pointers are indices into this same array), roots = roots*W words (the secondary array,
 * kmer_hash_table.cpp:2144-2145).  Either pointer may be NULL.  Sizes come from kg_compact_stats.           */
int kg_kaarme_download(kg_ctx* ctx, uint64_t* slots, uint64_t* roots);
/* The inverse: load a compact structure (e.g. one saved with kg_kaarme_download) into a KG_TABLE_KAARME context that
 * has not counted anything; kg_export / kg_export_text then decode it on the GPU (reconstruct_kmer_in_slot,
 * kmer_hash_table.cpp:3848-4058).  Malformed chains are detected while decoding (KG_ECUDA from the export), never
 * followed out of bounds.  Single GPU (world == 1).                                                            */
int kg_kaarme_upload(kg_ctx* ctx, const uint64_t* slots, uint64_t n_kmers, const uint64_t* roots, uint64_t n_roots);
/* Order-independent checksum of what kg_export would deliver (same selection: reported count >= min_abundance):
 * out[0] = k-mers, out[1] = sum of counts, out[2] = sum of g(k-mer), out[3] = sum of g(k-mer) * count (mod 2^64),
 * g = a 64-bit mix of the key words.  Computed on the device (plain table, or decoded from the compact structure).
 * The four words ADD over shards, so a sharded run can be compared with a single table without moving the k-mers --
 * the role sort + pytools/compare_outputs.py:1-33 plays for two output files.                                   */
int kg_checksum(kg_ctx* ctx, uint64_t min_abundance, int count_mode, uint64_t out[4]);
/* Table geometry for reports: bytes per slot and slots. */
int kg_table_info(const kg_ctx* ctx, uint64_t* slots, uint32_t* slot_bytes, uint32_t* key_words);

/* ---- roofline support -------------------------------------------------------------------------- */
/* Random 32-byte-sector atomic ceiling: n_ops uniform-random atomicAdd on
 * 32-byte-aligned slots over region_bytes; returns sectors/s (CUDA events, best of reps).           */
int kg_atomic_ceiling(int device, uint64_t region_bytes, uint64_t n_ops, int reps, double* sectors_per_s);
/* Number of kernel launches issued by this context so far (bench.py's gpu_launches). */
int kg_launch_count(const kg_ctx* ctx, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif
