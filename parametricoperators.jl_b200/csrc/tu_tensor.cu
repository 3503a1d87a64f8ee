// TU of the weight-streaming ParTensor kernels (po_tensor.cuh)
#define PO_TU_TENSOR
#include "po_tensor.cuh"
