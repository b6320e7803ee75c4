from diffuscene_b200.networks import (adjust_learning_rate, build_network, optimizer_factory,  # noqa: F401
                                      schedule_factory, LearningRateSchedule, StepLearningRateSchedule,
                                      LambdaLearningRateSchedule, WarmupCosineLearningRateSchedule)
from diffuscene_b200.networks.diffusion_scene_layout_ddpm import (DiffusionSceneLayout_DDPM,  # noqa: F401
                                                                   train_on_batch, validate_on_batch)
