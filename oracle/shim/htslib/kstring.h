/* Test scaffolding (oracle/_ref build only): minimal stand-in for <htslib/kstring.h>.
   The reference decoder classes only need the kstring_t POD and the two
   round-up macros; no htslib code is executed on the decode path. */
#ifndef PHQ_SHIM_KSTRING_H
#define PHQ_SHIM_KSTRING_H
#include <stddef.h>
#include <stdint.h>
#include <string.h>
typedef struct kstring_t { size_t l, m; char* s; } kstring_t;
#ifndef kroundup32
#define kroundup32(x) (--(x), (x)|=(x)>>1, (x)|=(x)>>2, (x)|=(x)>>4, (x)|=(x)>>8, (x)|=(x)>>16, ++(x))
#endif
#ifndef kroundup_size_t
#define kroundup_size_t(x) (--(x), (x)|=(x)>>(sizeof(size_t)/8), (x)|=(x)>>(sizeof(size_t)/4), (x)|=(x)>>(sizeof(size_t)/2), (x)|=(x)>>(sizeof(size_t)), (x)|=(x)>>(sizeof(size_t)*2), (x)|=(x)>>(sizeof(size_t)*4), ++(x))
#endif
#endif
