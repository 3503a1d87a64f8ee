// TU of the third-generation inverse kernel (po_fused3.cuh: staged-input pipelined inverse)
#define PO_TU_STAGE
#include "po_fused3.cuh"
