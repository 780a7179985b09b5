"""CPU oracle for the MAML inner-loop hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker or as the
timed CPU baseline.  The product package (``meta_interpolation_b200``) never
imports ``oracle`` and raises if its CUDA library is missing.

What is here
------------
* ``sepconv_op``   -- restatement of the reference's adaptive separable
  convolution kernels (reference ``sepconv/sepconv_op/sepconv.py:5-30,138-190``).
* ``backbones``    -- functional (params-dict) forward of the backbones on
  plain ATen CPU ops (reference ``sepconv/model.py:252-350`` ...).
* ``inner_rules``  -- the six inner-loop update rules
  (reference ``inner_loop_optimizers.py:136-244, 324-426``).
* ``maml``         -- the per-task inner loop + outer query pass
  (reference ``meta_learning_system.py:346-472, 584-606``).
* ``ops_ref``      -- per-operator NHWC references that mirror the C ABI in
  ``include/mi_b200.h`` one to one; the host-side executor is tested on CPU by
  injecting this table in place of the CUDA one.
* ``reference_shims`` + ``make_golden`` -- run the UNMODIFIED reference from
  ``/root/reference`` on CPU (container only) and write ``tests/golden/``.

Pinning
-------
The reference ships no tests and no golden vectors for this path
(SURVEY.md section 4).  The oracle is therefore pinned against outputs of the
reference itself, executed in the build container under the shims of
``reference_shims.py``; the resulting vectors are committed under
``tests/golden/`` with the generating script (``oracle/make_golden.py``).
"""
