"""vispec_b200: B200-native draft-and-verify decode path behind the API of KangJialiang/ViSpec's `vispec.model`.

    from vispec_b200.model.spec_model_ours import SpecModel        # (alias: EaModel)
    model.specgenerate(...)                                        # (alias: ea_generate)

The hot path is hand-written sm_100a CUDA in vispec_b200/csrc behind the C ABI of include/vispec_b200.h.
"""
__version__ = "0.1.0"
