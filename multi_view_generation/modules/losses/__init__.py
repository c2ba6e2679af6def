from multi_view_generation.modules.losses.vqperceptual import DummyLoss  # noqa: F401
