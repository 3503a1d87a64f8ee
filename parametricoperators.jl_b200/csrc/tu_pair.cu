// TU of the fourth-generation CTA-pair (cluster of 2) fused kernels (po_fused4.cuh)
#define PO_TU_PAIR
#include "po_fused4.cuh"
