/* test scaffolding: empty stand-in for <htslib/hts_endian.h>; only type names from sam.h/kstring.h are needed by the decoder classes */
