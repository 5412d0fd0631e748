"""Parity tests proper: the CUDA path (libpbrl.so through the C ABI) against the oracle on a B200."""
import pytest

import parity_cases as pc
from helpers import GpuBackend

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu_nograv(model_nograv):
    return GpuBackend(model_nograv)


@pytest.fixture(scope="module")
def gpu_grav(model_grav):
    return GpuBackend(model_grav)


def test_forward_stages_limits(gpu_nograv, model_nograv):
    pc.check_forward_stages(gpu_nograv, model_nograv, gravity=False, N=64)


def test_forward_stages_contacts(gpu_grav, model_grav):
    pc.check_forward_stages(gpu_grav, model_grav, gravity=True, N=64)


def test_one_substep(gpu_nograv, model_nograv, gpu_grav, model_grav):
    pc.check_one_substep(gpu_nograv, model_nograv, gravity=False, N=256)
    pc.check_one_substep(gpu_grav, model_grav, gravity=True, N=256)


@pytest.mark.parametrize("pd_mode", [0, 1, 2, 3])
def test_32_contact_free_substeps(gpu_nograv, model_nograv, pd_mode):
    pc.check_32_substeps(gpu_nograv, model_nograv, pd_mode=pd_mode, N=128)


def test_pd_laws(gpu_nograv, model_nograv):
    pc.check_pd_torque(gpu_nograv, model_nograv, N=64)


def test_golden_T1(gpu_nograv, model_nograv):
    pc.check_golden_T1(gpu_nograv, model_nograv)


def test_golden_T2_T4(gpu_nograv, model_nograv):
    pc.check_golden_T2_T4(gpu_nograv, model_nograv)


def test_replay(gpu_nograv, model_nograv):
    pc.check_replay(gpu_nograv, model_nograv, N=4096)


def test_reset(gpu_nograv, model_nograv):
    pc.check_reset(gpu_nograv, model_nograv)


def test_hands_against_thighs(gpu_nograv, model_nograv):
    """Body-vs-body contacts (sphere / capsule candidates) on the CUDA path, dense Hessian branch included."""
    pc.check_hands_against_thighs(gpu_nograv, model_nograv, N=48, steps=12)


def test_contact_rollout(gpu_grav, model_grav):
    pc.check_contact_rollout(gpu_grav, model_grav, N=64, steps=40)


def test_mimic_mode(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    b = GpuBackend(model_grav, mimic_desc=mdesc)
    pc.check_mimic(b, model_grav, table, mdesc, N=112)


def test_autoreset(model_grav):
    table, mdesc, keep = pc.mimic_setup(model_grav)
    pc.check_autoreset(GpuBackend(model_grav, mimic_desc=mdesc), model_grav, table, mdesc, N=133)
