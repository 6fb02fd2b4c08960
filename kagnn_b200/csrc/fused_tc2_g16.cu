// The pipelined kernel with the gather-heavy role split: 8 basis-producer warps + 16 gather warps (see the note above
// KAGNN_TC2_ENTRY in fused_tc2.cu).  Same source, second instantiation: exports kagnn_fused_fwd_tc2_g16 only.
#define KAGNN_TC2_NPW 8
#define KAGNN_TC2_NGW 16
#define KAGNN_TC2_VARIANT_G16 1
#include "fused_tc2.cu"
