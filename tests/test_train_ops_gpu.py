"""Parity of the training-step kernels (weight gradient on tcgen05, input gradient through the forward kernel,
train-mode BatchNorm forward/backward, max-pool backward, deformable sampling forward/backward) against torch
autograd in fp32 on the CPU.  Tolerances are stated in train_cases."""
import pytest

import train_cases as TC


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(TC.WGRAD_CASES))
def test_wgrad_case(cuda_lib, name):
    TC.run_wgrad_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(TC.DGRAD_CASES))
def test_dgrad_case(cuda_lib, name):
    TC.run_dgrad_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(residual=False), dict(relu=False, residual=False, Cc=64, rows=777),
                                dict(Cc=2048, rows=600), dict(dtype="f16")])
def test_batchnorm_train(cuda_lib, kw):
    TC.run_bn_case(**kw)


@pytest.mark.gpu
def test_maxpool_bwd(cuda_lib):
    TC.run_maxpool_bwd_case()


@pytest.mark.gpu
def test_gradient_joins(cuda_lib):
    TC.run_add_cases()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(TC.DCN_CASES))
def test_dcn_train(cuda_lib, name):
    TC.run_dcn_case(name)


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(Cc=128, HW=777, B=2)])
def test_groupnorm_bwd(cuda_lib, kw):
    TC.run_gn_case(**kw)


@pytest.mark.gpu
def test_resample_bwd(cuda_lib):
    TC.run_resample_bwd_cases()


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(up=1), dict(up=2), dict(up=2, accumulate=True), dict(up=1, H=3, W=3)])
def test_reflect_dgrad(cuda_lib, kw):
    TC.run_reflect_dgrad_case(**kw)


@pytest.mark.gpu
def test_softplus_bwd(cuda_lib):
    TC.run_softplus_case()


@pytest.mark.gpu
def test_weight_pack_kernels(cuda_lib):
    TC.run_pack_cases()
