"""Parity of the tcgen05 implicit-GEMM convolution (called through the C ABI) against fp32 torch CPU
operators on identical 16-bit-rounded operands.  Tolerances are stated in conv_cases.run_case."""
import pytest

import conv_cases


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(conv_cases.CASES))
def test_conv_case(cuda_lib, name):
    conv_cases.run_case(name)
