/* Test scaffolding (oracle/_ref build only): minimal stand-in for <htslib/sam.h>.
   Declares just the types and flag constants that appear in the signatures of
   the reference headers pulled in by the decoder classes. None of the inline
   stubs below is reached by PamlDecoder / MdDecoder / NaiveMolecularDecoder. */
#ifndef PHQ_SHIM_SAM_H
#define PHQ_SHIM_SAM_H
#include <stdint.h>
#include <stddef.h>
typedef int64_t hts_pos_t;
enum htsFormatCategory { unknown_category, sequence_data, variant_data, index_file, region_list, category_maximum = 32767 };
enum htsExactFormat { unknown_format, binary_format, text_format, sam, bam, bai, cram, crai, vcf, bcf, csi, gzi, tbi, bed, htsget, json, empty_format, fasta_format, fastq_format, fai_format, fqi_format, format_maximum = 32767 };
enum htsCompression { no_compression, gzip, bgzf, custom, bzip2_compression, razf_compression, xz_compression, zstd_compression, compression_maximum = 32767 };
typedef struct htsFormat { enum htsFormatCategory category; enum htsExactFormat format; struct { short major, minor; } version; enum htsCompression compression; short compression_level; void* specific; } htsFormat;
typedef struct bam_hdr_t { int32_t n_targets, ignore_sam_err; size_t l_text; uint32_t* target_len; char** target_name; char* text; void* sdict; void* hrecs; uint32_t ref_count; } bam_hdr_t; /* the reference's include.h adds: typedef bam_hdr_t sam_hdr_t */
typedef struct bam1_core_t { hts_pos_t pos; int32_t tid; uint16_t bin; uint8_t qual; uint8_t l_extranul; uint16_t flag; uint16_t l_qname; uint32_t n_cigar; int32_t l_qseq; int32_t mtid; hts_pos_t mpos; hts_pos_t isize; } bam1_core_t;
typedef struct bam1_t { bam1_core_t core; uint64_t id; uint8_t* data; int l_data; uint32_t m_data; uint32_t mempolicy; } bam1_t;
typedef struct htsThreadPool { void* pool; int qsize; } htsThreadPool;
typedef struct hFILE hFILE;
typedef struct htsFile htsFile;
typedef struct BGZF BGZF;
#define BAM_FPAIRED 1
#define BAM_FPROPER_PAIR 2
#define BAM_FUNMAP 4
#define BAM_FMUNMAP 8
#define BAM_FREVERSE 16
#define BAM_FMREVERSE 32
#define BAM_FREAD1 64
#define BAM_FREAD2 128
#define BAM_FSECONDARY 256
#define BAM_FQCFAIL 512
#define BAM_FDUP 1024
#define BAM_FSUPPLEMENTARY 2048
#define bam_get_qname(b) ((char*)(b)->data)
#define bam_get_aux(b) ((b)->data)
#define bam_get_l_aux(b) (0)
static inline uint32_t le_to_u32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static inline uint16_t le_to_u16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline int64_t bam_aux2i(const uint8_t*) { return 0; }
static inline double bam_aux2f(const uint8_t*) { return 0; }
static inline char bam_aux2A(const uint8_t*) { return 0; }
static inline char* bam_aux2Z(const uint8_t*) { return 0; }
static inline int bam_aux_append(bam1_t*, const char*, char, int, const uint8_t*) { return 0; }
#endif
