from .wrapper import PipelineParallelWrapper
from .trainer import PipelineTrainer
from .dataloader import PipelineDataLoader
from .schedule import PipelineSchedule, AllFwdAllBwdSchedule, OneFOneBSchedule
from .utils import distribute_layers

__all__ = ["PipelineParallelWrapper", "PipelineTrainer", "PipelineDataLoader", "PipelineSchedule",
           "AllFwdAllBwdSchedule", "OneFOneBSchedule", "distribute_layers"]
