"""Drop-in import paths of the reference package (alexanderswerdlow/BEVGen) for the two hot paths.

Hydra `_target_` strings such as `multi_view_generation.modules.stage1.vqgan.VQModel` keep resolving; the classes
here hold parameters under the reference's state-dict key names and route all arithmetic to `bevgen_b200`
(sm_100a kernels through the C ABI).  Nothing else of the reference (data loading, training, logging) is provided.
"""
