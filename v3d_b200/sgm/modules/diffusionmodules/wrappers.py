from v3d_b200.sampling import IdentityWrapper, OpenAIWrapper  # noqa: F401  (reference: wrappers.py:9-34)

OPENAIUNETWRAPPER = "v3d_b200.sgm.modules.diffusionmodules.wrappers.OpenAIWrapper"
