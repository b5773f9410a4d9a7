"""The reference's OWN tests and tutorials, unmodified, on the plug-in classes.

``pySDC/tests/test_tutorials/test_step_{1..6}.py`` (tutorial steps 1-6 with the asserts pySDC ships: spatial and
collocation accuracy, iteration counts of SDC / MLSDC / PFASST, ...), ``tests/test_transfer_classes/test_mesh_to_mesh.py``,
``tests/test_2d_fd_accuracy.py`` and ``tests/test_convergence_controllers/test_check_convergence.py`` are imported from the reference tree (or its shipped copy ``oracle/_ref``) and their
test functions are called as they are.  The only change is WHERE the class names resolve: for the duration of a test the
modules ``pySDC.implementations.problem_classes.{HeatEquation_ND_FD, AdvectionEquation_ND_FD, AllenCahn_2D_FD}``,
``sweeper_classes.{generic_implicit, imex_1st_order, multi_implicit}``, ``transfer_classes.TransferMesh`` and
``datatype_classes.mesh`` are replaced by modules that export the classes of ``pysdc_b200.pysdc_plugin`` — the two
changed import lines of INTEGRATION.md, applied to the reference's own test-suite.  Controller, Step / Level,
collocation, hooks, statistics and the tutorials' own code stay the reference's.

``numpy`` variants: kernel library replaced by the numpy test double (CPU suite); ``cuda`` variants (``-m gpu``): the
real CUDA kernels.  Not run: tutorial tests that need classes outside the path (Penning trap: step 3 B/C, 4 D), and the
``mpirun`` launcher (6 C).  Steps 5 B / C (PFASST; C: advection with ``solver_type='direct'`` = GMRES to 1e-14 here) take a
minute each on the numpy double and run in the GPU suite only (2 s / 7 s there); tests/test_pysdc_dropin.py keeps a PFASST run
under the reference's controller in the CPU suite.
``matplotlib`` is absent from the image; the tutorials only plot with it, so a do-nothing stand-in is installed."""
import importlib
import os
import sys
import types
from unittest import mock

import pytest

from conftest import reference_paths

REF_PATHS = reference_paths()
pytestmark = pytest.mark.skipif(REF_PATHS is None, reason="reference tree not present (no /root/reference, no oracle/_ref)")

SWAPPED = {
    "pySDC.implementations.problem_classes.HeatEquation_ND_FD": ["heatNd_unforced", "heatNd_forced"],
    "pySDC.implementations.problem_classes.AdvectionEquation_ND_FD": ["advectionNd"],
    "pySDC.implementations.problem_classes.AllenCahn_2D_FD": [
        "allencahn_fullyimplicit", "allencahn_semiimplicit", "allencahn_semiimplicit_v2", "allencahn_multiimplicit",
        "allencahn_multiimplicit_v2"],
    "pySDC.implementations.sweeper_classes.generic_implicit": ["generic_implicit"],
    "pySDC.implementations.sweeper_classes.imex_1st_order": ["imex_1st_order"],
    "pySDC.implementations.sweeper_classes.multi_implicit": ["multi_implicit"],
    "pySDC.implementations.transfer_classes.TransferMesh": ["mesh_to_mesh"],
    "pySDC.implementations.datatype_classes.mesh": ["mesh", "imex_mesh", "comp2_mesh"],
}

REFERENCE_TESTS = [
    ("pySDC.tests.test_tutorials.test_step_1", "test_A"), ("pySDC.tests.test_tutorials.test_step_1", "test_B"),
    ("pySDC.tests.test_tutorials.test_step_1", "test_C"), ("pySDC.tests.test_tutorials.test_step_1", "test_D"),
    ("pySDC.tests.test_tutorials.test_step_2", "test_A"), ("pySDC.tests.test_tutorials.test_step_2", "test_B"),
    ("pySDC.tests.test_tutorials.test_step_2", "test_C"), ("pySDC.tests.test_tutorials.test_step_3", "test_A"),
    ("pySDC.tests.test_tutorials.test_step_4", "test_A"), ("pySDC.tests.test_tutorials.test_step_4", "test_B"),
    ("pySDC.tests.test_tutorials.test_step_4", "test_C"), ("pySDC.tests.test_tutorials.test_step_5", "test_A"),
    ("pySDC.tests.test_tutorials.test_step_5", "test_B", {}, "gpu-only"),
    ("pySDC.tests.test_tutorials.test_step_5", "test_C", {}, "gpu-only"),
    ("pySDC.tests.test_tutorials.test_step_6", "test_A"),
    ("pySDC.tests.test_tutorials.test_step_6", "test_B"),
    ("pySDC.tests.test_transfer_classes.test_mesh_to_mesh", "test_mesh_to_mesh_1d_dirichlet"),
    ("pySDC.tests.test_transfer_classes.test_mesh_to_mesh", "test_mesh_to_mesh_1d_periodic"),
    ("pySDC.tests.test_transfer_classes.test_mesh_to_mesh", "test_mesh_to_mesh_2d_periodic"),
    ("pySDC.tests.test_2d_fd_accuracy", "test_spatial_accuracy"),
    # convergence_controller_classes/check_convergence.py on a 1-D periodic order-6 heat problem with the 'direct' solver
    # (parametrised in the reference: the arguments are its own parameter lists; "gpu-only": thousands of CG iterations
    # per solve stand in for the 'direct' solver, which the numpy test double takes minutes for)
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_iter", dict(maxiter=1)),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_iter", dict(maxiter=5), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_iter", dict(maxiter=50), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_increment", dict(e_tol=1e-3), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_increment", dict(e_tol=1e-5), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_increment", dict(e_tol=1e-10), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_residual", dict(restol=1e-3)),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_residual", dict(restol=1e-5), "gpu-only"),
    ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_residual", dict(restol=1e-10), "gpu-only"),
]
# on the GPU: one test per kind of run (collocation set-up, SDC through the front end, MLSDC, PFASST, transfer orders)
ON_GPU = {("pySDC.tests.test_tutorials.test_step_1", "test_B"), ("pySDC.tests.test_tutorials.test_step_2", "test_C"),
          ("pySDC.tests.test_tutorials.test_step_3", "test_A"), ("pySDC.tests.test_tutorials.test_step_4", "test_C"),
          ("pySDC.tests.test_tutorials.test_step_5", "test_B"), ("pySDC.tests.test_tutorials.test_step_6", "test_A"),
          ("pySDC.tests.test_transfer_classes.test_mesh_to_mesh", "test_mesh_to_mesh_2d_periodic"),
          ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_increment"),
          ("pySDC.tests.test_convergence_controllers.test_check_convergence", "test_convergence_by_residual")}


SETUP_ONLY = {("pySDC.tests.test_tutorials.test_step_1", "test_C"), ("pySDC.tests.test_tutorials.test_step_1", "test_D"),
              ("pySDC.tests.test_tutorials.test_step_4", "test_B"), ("pySDC.tests.test_tutorials.test_step_5", "test_A")}


def _is_scratch(name):
    return name.startswith(("pySDC.tutorial", "pySDC.tests")) or name == "matplotlib" or name.startswith("matplotlib.")


@pytest.fixture
def swapped_reference(request, tmp_path, monkeypatch):
    """The reference importable, its modules of the path replaced by plug-in exports, cwd = a scratch directory (the
    tutorials write their output to ./data); everything is undone afterwards so that other tests see the real modules."""
    for p in reversed(REF_PATHS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pysdc_b200 import backend

    old_backend = backend._backend
    if request.param == "cuda":
        backend.set_backend(backend.CudaBackend())
    else:
        from fake_backend import NumpyBackend

        backend.set_backend(NumpyBackend())
    from pysdc_b200 import pysdc_plugin as plugin

    saved = {k: sys.modules.get(k) for k in SWAPPED}
    stale = [k for k in sys.modules if _is_scratch(k)]
    for k in stale:  # tutorial modules bind the class names at import time: they have to be imported afresh
        del sys.modules[k]
    for name, exports in SWAPPED.items():
        mod = types.ModuleType(name)
        mod.__doc__ = "plug-in classes of pysdc_b200 under the reference's module name (tests/test_reference_suite.py)"
        for e in exports:
            setattr(mod, e, getattr(plugin, e))
        sys.modules[name] = mod
    if importlib.util.find_spec("matplotlib") is None:
        fake = mock.MagicMock(name="matplotlib")
        for plt in (fake.pyplot, fake.pylab):  # the tutorials assert that the figure file exists
            plt.savefig.side_effect = lambda fname, *a, **k: open(fname, "wb").close()
        sys.modules.update({"matplotlib": fake, "matplotlib.pyplot": fake.pyplot, "matplotlib.pylab": fake.pylab})
    monkeypatch.chdir(tmp_path)
    os.makedirs(os.path.join(tmp_path, "data"), exist_ok=True)
    try:
        yield plugin
    finally:
        for k in [k for k in sys.modules if _is_scratch(k)]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        backend.set_backend(old_backend)


def _params():
    out = []
    for mod, fn, *rest in REFERENCE_TESTS:
        kwargs = rest[0] if rest else {}
        tag = f"{mod.split('.')[-1]}::{fn}" + "".join(f"-{v}" for v in kwargs.values())
        if "gpu-only" not in rest:
            out.append(pytest.param("numpy", mod, fn, kwargs, id=f"numpy-{tag}"))
        if (mod, fn) in ON_GPU or "gpu-only" in rest:
            out.append(pytest.param("cuda", mod, fn, kwargs, id=f"cuda-{tag}", marks=pytest.mark.gpu))
    return out


@pytest.mark.parametrize("swapped_reference,module,function,kwargs", _params(), indirect=["swapped_reference"])
def test_reference_test_passes_on_plugin_classes(swapped_reference, module, function, kwargs):
    from pysdc_b200 import backend

    launches0 = backend.get_backend().launches
    getattr(importlib.import_module(module), function)(**kwargs)
    # the test really went through the device classes (set-up-only tutorials launch nothing: u_exact is a host expression)
    assert backend.get_backend().launches > launches0 or (module, function) in SETUP_ONLY


@pytest.mark.parametrize("kind", ["numpy", pytest.param("cuda", marks=pytest.mark.gpu)])
def test_reference_cpu_vs_gpu_check_of_the_heat_class(kind, tmp_path, monkeypatch):
    """pySDC/projects/GPU/heat.py::main, unmodified: the reference runs the same IMEX-SDC description (3-D periodic forced
    heat, 32^3, CG) once with its CPU class and once with its GPU class (``HeatEquation_ND_FD_CuPy.heatNd_forced``) and
    asserts ``abs(uend_gpu.get() - uend_cpu) < 1e-13`` (heat.py:94).  Here only the GPU class's module resolves to the
    plug-in; the CPU class, the sweeper and the controller are the reference's."""
    for p in reversed(REF_PATHS):
        if p not in sys.path:
            sys.path.insert(0, p)
    from pysdc_b200 import backend

    old_backend = backend._backend
    if kind == "cuda":
        backend.set_backend(backend.CudaBackend())
    else:
        from fake_backend import NumpyBackend

        backend.set_backend(NumpyBackend())
    from pysdc_b200 import pysdc_plugin as plugin

    name = "pySDC.implementations.problem_classes.HeatEquation_ND_FD_CuPy"
    saved = sys.modules.get(name)
    mod = types.ModuleType(name)
    mod.heatNd_forced, mod.heatNd_unforced = plugin.heatNd_forced, plugin.heatNd_unforced
    sys.modules[name] = mod
    sys.modules.pop("pySDC.projects.GPU.heat", None)
    monkeypatch.chdir(tmp_path)
    try:
        launches0 = backend.get_backend().launches
        importlib.import_module("pySDC.projects.GPU.heat").main()
        assert backend.get_backend().launches > launches0
    finally:
        sys.modules.pop("pySDC.projects.GPU.heat", None)
        if saved is None:
            sys.modules.pop(name, None)
        else:
            sys.modules[name] = saved
        backend.set_backend(old_backend)
