"""meta_interpolation_b200 -- B200-native MAML inner-loop hot path of myungsub/meta-interpolation.

Public surface mirrors the reference's modules for this path:

* ``meta_learning_system.SceneAdaptiveInterpolation``  (reference meta_learning_system.py:29)
* ``inner_loop_optimizers.LSLRGradientDescentLearningRule`` / ``MetaSGDLearningRule``
  (reference inner_loop_optimizers.py:57, :248)
* ``sepconv.model.MetaNetwork`` and ``sepconv.sepconv_op.sepconv.FunctionSepconv``
  (reference sepconv/model.py:168, sepconv/sepconv_op/sepconv.py:247)
* ``model_utils.extract_top_level_dict``               (reference model_utils.py:272)

All compute is hand-written sm_100a CUDA behind the C ABI in ``include/mi_b200.h``
(``libmi_b200.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
