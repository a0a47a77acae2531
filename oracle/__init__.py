"""CPU oracle for the Ferrite.jl global-assembly hot path.

TEST INFRASTRUCTURE ONLY.  This package is a plain numpy (+ one small C file)
restatement of the reference algorithm (Ferrite.jl v1.6 "Next", pure Julia).  It
exists so that the CUDA library `libferrite_b200.so` can be checked for parity.
Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import it.  The product path
(`ferrite.jl_b200/`) never imports it and has no CPU fallback.

Parity pinning: the reference cannot be executed here (no Julia toolchain in the
image, no network), so the oracle is pinned to the reference's own golden
values instead -- see `tests/test_oracle_goldens.py`:
  * norm(K.nzval) = 1138.8803468514259   docs/src/topics/assembly.md:348-356
  * heat tutorial norm(u) = 3.307743912641305   docs/src/literate-tutorials/heat_equation.jl:234
  * literal celldofs vectors   test/test_dofs.jl:71-126,256-257
  * literal prescribed_dofs / inhomogeneities   test/test_constraints.jl:100-101,171-172, test/test_dofs.jl:257
  * 3-dof assemble!+apply! KAT   test/test_assembler_extensions.jl:69-86
  * hyperelasticity tutorial norm(u) = 4.761404305083876   docs/src/literate-tutorials/hyperelasticity.jl:442
    (Neo-Hooke tangent/residual + facet traction + inhomogeneous Dirichlet + apply_zero!)
Third-party arithmetic not vendored under the reference tree: Tensors.jl 1.17.1
(det/inv/otimes closed forms, restated here), ForwardDiff 1.4.1 (shape-function
gradients; restated as analytic derivatives), SparseArrays 1.12 (CSC layout).
"""
from .refshapes import *      # noqa: F401,F403
from .interpolations import *  # noqa: F401,F403
from .quadrature import *     # noqa: F401,F403
from .grid import *           # noqa: F401,F403
from .dofs import *           # noqa: F401,F403
from .pattern import *        # noqa: F401,F403
from .element import *        # noqa: F401,F403
from .assemble import *       # noqa: F401,F403
from .constraints import *    # noqa: F401,F403
from .facets import *         # noqa: F401,F403
from .affine import *        # noqa: F401,F403
from .mixed import *         # noqa: F401,F403
