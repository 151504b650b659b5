"""``PipelineDataLoader`` — infinite micro-batch iterator
(reference ``parallelism/pipeline_parallel/dataloader.py:2-56``).
``grad_acc_steps`` is the number of micro-batches per optimizer step."""
from __future__ import annotations


class PipelineDataLoader:
    def __init__(self, dataloader, grad_acc_steps: int, task_type: str = "auto"):
        self.dataloader = dataloader
        self.grad_acc_steps = int(grad_acc_steps)
        self.task_type = task_type
        self.data_iter = iter(dataloader)
        self.epoch = 0

    @property
    def iterator(self):               # the reference's attribute name for the live iterator
        return self.data_iter

    def __iter__(self):
        return self

    def __len__(self):
        return len(self.dataloader)

    def _next_raw(self):
        try:
            return next(self.data_iter)
        except StopIteration:
            self.epoch += 1
            sampler = getattr(self.dataloader, "sampler", None)
            if hasattr(sampler, "set_epoch"):
                sampler.set_epoch(self.epoch)
            self.data_iter = iter(self.dataloader)
            return next(self.data_iter)

    def __next__(self):
        batch = self._next_raw()
        if self.task_type == "auto":
            if isinstance(batch, dict) and "input_ids" in batch:
                self.task_type = "clm"
            else:
                self.task_type = "classification"
        if self.task_type == "clm":
            return batch
        if isinstance(batch, dict):
            if "images" in batch:
                return batch
            return {"images": batch["image"], "labels": batch["label"]}
        images, labels = batch
        return {"images": images, "labels": labels}
