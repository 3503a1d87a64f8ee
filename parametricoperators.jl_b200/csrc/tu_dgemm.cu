// TU of the FP64 GEMM arms for large dense mode products (po_dgemm.cuh)
#define PO_TU_DGEMM
#include "po_dgemm.cuh"
