// TU of the second-generation fused (t, x) kernels (po_fused2.cuh): pipelined and single-tile variants
#define PO_TU_FUSED2
#include "po_fused2.cuh"
