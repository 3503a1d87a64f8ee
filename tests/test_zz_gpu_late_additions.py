"""GPU wrappers that were added AFTER the round's GPU budget was spent (-m gpu): new PLAN shapes over kernels that are all GPU-verified
(the widening pass compiles few-mode layers to the existing fused / pruned-c16 / gather kernels; the reference's test/transforms.jl
expressions use the generic matrix / DFT / diagonal kernels).  The identical test bodies pass on the CPU kernel emulator, under
AddressSanitizer and ThreadSanitizer (profiles/r2q_emulator_sanitizers.txt).  The file name sorts last so that the GPU-verified suite runs first."""
import pytest

from tests import test_emul_parity as E

pytestmark = pytest.mark.gpu


def test_reference_transforms_jl_expressions(cuda_engine):
    E.test_reference_transforms_jl_expressions(cuda_engine)


@pytest.mark.parametrize("shape,c,kx,kt,batch", [([4, 64, 16], 4, 4, 5, 2), ([32, 64, 32], 6, 8, 4, 1),
                                                 ([32, 32, 64, 16], 20, 4, 4, 1)])  # last: >= 4 x 148 units -> row ring / staged inverse / c16
def test_fewer_kept_modes_reach_the_fused_kernels(cuda_engine, shape, c, kx, kt, batch):
    E.test_fewer_kept_modes_reach_the_fused_kernels(cuda_engine, shape, c, kx, kt, batch)
