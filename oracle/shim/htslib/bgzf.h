/* test scaffolding: empty stand-in for <htslib/bgzf.h>; only type names from sam.h/kstring.h are needed by the decoder classes */
