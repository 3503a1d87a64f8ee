// TU of the third-generation forward kernel (po_fused3.cuh: asynchronous row ring)
#define PO_TU_RING
#include "po_fused3.cuh"
