"""Drop-in for mars/utils/evaluation.py: prediction-vs-ground-truth areas on the device."""
import torch

from . import ops


class Evaluator:
    r"""Computes intersection and union between prediction and ground-truth (Evaluator.classify_prediction)."""

    @classmethod
    def initialize(cls):
        cls.ignore_index = 255

    @classmethod
    def classify_prediction(cls, pred_mask: torch.Tensor, batch: dict):
        """pred_mask [B,H,W]; batch['query_mask'] [B,H,W]; optional batch['query_ignore_idx'].

        Returns (area_inter [2,B], area_union [2,B]) float tensors like the reference (mars/utils/evaluation.py:12-38);
        unlike the reference the inputs are not modified in place.
        """
        gt = batch.get("query_mask")
        ignore = batch.get("query_ignore_idx")
        if ignore is not None:
            assert torch.logical_and(ignore, gt).sum() == 0
        dev = pred_mask.device if pred_mask.is_cuda else torch.device("cuda")
        out = ops.eval_areas(pred_mask.to(dev).float(), gt.to(dev).float(),
                             None if ignore is None else ignore.to(dev).float())
        out = out.float()
        return out[:, :2].t().contiguous(), out[:, 2:].t().contiguous()


NCLASS = {"pascal": 20, "pascal5i": 20, "coco": 80, "fss": 1000, "paco_part": 448, "pascal_part": 100, "lvis": 1203}


class AverageMeter:
    r"""Stores evaluation results (drop-in for mars/utils/logger.py:14-100) with the buffers on the device.

    `intersection_buf` / `union_buf` are exact int64 pixel counts [2, nclass] (the reference keeps float32);
    `update` takes the (area_inter, area_union) pair `Evaluator.classify_prediction` returns, or the raw int32
    areas via `update_areas`, without any host synchronisation.  `all_reduce` sums the buffers over the ranks of a
    data-parallel evaluation (episodes are sharded).
    """

    def __init__(self, dataset, device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("marsb200.AverageMeter keeps its buffers on a CUDA device (no CPU path)")
        self.benchmark = dataset.benchmark
        ids = list(dataset.class_ids)
        if self.benchmark == "pascal5i":
            ids = [i - 1 for i in ids]  # 1-index to 0-index (logger.py:21-23)
        self.class_ids_interest = torch.tensor(ids, dtype=torch.int64, device=self.device)
        self.nclass = NCLASS[self.benchmark]
        self.intersection_buf = torch.zeros((2, self.nclass), dtype=torch.int64, device=self.device)
        self.union_buf = torch.zeros((2, self.nclass), dtype=torch.int64, device=self.device)
        self.class_ids_known_bad = []
        self.intersection_buf_known_bad = torch.zeros_like(self.intersection_buf)
        self.union_buf_known_bad = torch.zeros_like(self.union_buf)
        self.loss_buf, self.loss_buf_known_bad = [], []

    @staticmethod
    def _areas(inter_b, union_b):
        return torch.cat([inter_b.t(), union_b.t()], dim=1).round().to(torch.int32).contiguous()

    def update_areas(self, areas, class_id):
        """areas int32 [n, 4] straight from `ops.eval_areas`."""
        ops.eval_accumulate(areas, class_id.to(self.device), self.intersection_buf, self.union_buf)

    def update(self, inter_b, union_b, class_id, loss):
        self.update_areas(self._areas(inter_b.to(self.device), union_b.to(self.device)), class_id)
        self.loss_buf.append(torch.tensor(0.0) if loss is None else loss)

    def update_bad_preds(self, inter_b, union_b, class_id, loss):
        for c in class_id.reshape(-1).tolist():
            if c not in self.class_ids_known_bad:
                self.class_ids_known_bad.append(c)
        ops.eval_accumulate(self._areas(inter_b.to(self.device), union_b.to(self.device)), class_id.to(self.device),
                            self.intersection_buf_known_bad, self.union_buf_known_bad)
        self.loss_buf_known_bad.append(torch.tensor(0.0) if loss is None else loss)

    def _iou(self, inter, union, interest):
        out = ops.eval_iou(inter, union, interest)
        return out[0], out[1], out[2:2 + min(interest.numel(), 20)]

    def compute_iou(self):
        """-> (mIoU, FB-IoU, fg IoU of the first 20 classes of interest), logger.py:69-78."""
        return self._iou(self.intersection_buf, self.union_buf, self.class_ids_interest)

    def compute_iou_bad_preds(self):
        interest = torch.tensor(self.class_ids_known_bad, dtype=torch.int64, device=self.device)
        return self._iou(self.intersection_buf_known_bad, self.union_buf_known_bad, interest)

    def all_reduce(self, group=None):
        import torch.distributed as dist

        for buf in (self.intersection_buf, self.union_buf, self.intersection_buf_known_bad, self.union_buf_known_bad):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
