"""`TetrahedraNerfPipeline` -- the nerfstudio pipeline shell that hands the dataparser transform / scale and the
dataset metadata to the model (reference tetranerf/nerfstudio/pipeline.py:16-58).  Control plane, outside the
hot-path scope; it only exists so that `ns-train tetra-nerf` resolves.  Requires nerfstudio."""
import typing

from nerfstudio.pipelines.base_pipeline import DDP, Model, Pipeline, VanillaPipeline, VanillaPipelineConfig, dist


class TetrahedraNerfPipeline(VanillaPipeline):
    def __init__(self, config: VanillaPipelineConfig, device: str, test_mode: str = "val", world_size: int = 1, local_rank: int = 0,
                 grad_scaler=None):
        Pipeline.__init__(self)
        self.config, self.test_mode = config, test_mode
        self.datamanager = config.datamanager.setup(device=device, test_mode=test_mode, world_size=world_size, local_rank=local_rank)
        self.datamanager.to(device)
        assert self.datamanager.train_dataset is not None, "Missing input dataset"
        outputs = self.datamanager.train_dataparser_outputs
        extra = {"grad_scaler": grad_scaler} if grad_scaler is not None else {}
        self._model = config.model.setup(
            scene_box=self.datamanager.train_dataset.scene_box, num_train_data=len(self.datamanager.train_dataset),
            metadata=self.datamanager.train_dataset.metadata,
            dataparser_transform=outputs.dataparser_transform, dataparser_scale=outputs.dataparser_scale, **extra)
        self.model.to(device)
        self.world_size = world_size
        if world_size > 1:  # data-parallel over rays, one process per GPU (reference :53-58)
            self._model = typing.cast(Model, DDP(self._model, device_ids=[local_rank], find_unused_parameters=True))
            dist.barrier(device_ids=[local_rank])
