"""Model base classes.  Mirror of the training-relevant part of models/base_model.py: BaseModel
(optimizer/scheduler, :28-36) and MatchingBaseModel.compute_loss (:122-164).  The tracking evaluation loop
(:44-117, :166-247) needs the dataset stack (nuscenes-devkit, shapely, pyquaternion) and is out of scope for
the hot path."""
import torch
import torch.nn.functional as F

from ..compat import EasyDict, LightningModule


class BaseModel(LightningModule):
    def __init__(self, config=None, **kwargs):
        super().__init__()
        if config is None:
            config = EasyDict(kwargs)
        self.config = config

    def configure_optimizers(self):
        if self.config.optimizer.lower() == 'sgd':
            optimizer = torch.optim.SGD(self.parameters(), lr=self.config.lr, momentum=0.9,
                                        weight_decay=self.config.wd)
        else:
            optimizer = torch.optim.Adam(self.parameters(), lr=self.config.lr, weight_decay=self.config.wd,
                                         betas=(0.5, 0.999), eps=1e-06)
        scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=self.config.lr_decay_step,
                                                    gamma=self.config.lr_decay_rate)
        return {"optimizer": optimizer, "lr_scheduler": scheduler}

    def compute_loss(self, data, output):
        raise NotImplementedError

    def evaluate_one_sequence(self, sequence):
        raise NotImplementedError("tracking evaluation needs the dataset stack; out of scope for the hot path")


class MatchingBaseModel(BaseModel):
    def compute_loss(self, data, output):
        """Segmentation BCE, masked smooth-L1 vote loss, objectness BCE (pos_weight 2, thresholds 0.3 / 0.6)
        and masked smooth-L1 box loss — device-aware (the reference hard-codes `.cuda()`, base_model.py:151)."""
        estimation_boxes = output['estimation_boxes']        # (B, num_proposal, 5)
        estimation_cla = output['estimation_cla']            # (B, N)
        seg_label, box_label = data['seg_label'], data['box_label']
        proposal_center, vote_xyz = output["center_xyz"], output["vote_xyz"]

        loss_seg = F.binary_cross_entropy_with_logits(estimation_cla, seg_label)

        loss_vote = F.smooth_l1_loss(vote_xyz, box_label[:, None, :3].expand_as(vote_xyz), reduction='none')
        loss_vote = (loss_vote.mean(2) * seg_label).sum() / (seg_label.sum() + 1e-06)

        dist = torch.sqrt(torch.sum((proposal_center - box_label[:, None, :3]) ** 2, dim=-1) + 1e-6)
        objectness_label = (dist < 0.3).float()
        objectness_mask = ((dist < 0.3) | (dist > 0.6)).float()
        loss_objective = F.binary_cross_entropy_with_logits(
            estimation_boxes[:, :, 4], objectness_label, reduction='none',
            pos_weight=torch.full((1,), 2.0, device=estimation_boxes.device))   # device-side fill: graph-capturable
        loss_objective = torch.sum(loss_objective * objectness_mask) / (torch.sum(objectness_mask) + 1e-6)

        loss_box = F.smooth_l1_loss(estimation_boxes[:, :, :4],
                                    box_label[:, None, :4].expand_as(estimation_boxes[:, :, :4]), reduction='none')
        loss_box = torch.sum(loss_box.mean(2) * objectness_label) / (objectness_label.sum() + 1e-6)
        return {"loss_objective": loss_objective, "loss_box": loss_box, "loss_seg": loss_seg, "loss_vote": loss_vote}


class MotionBaseModel(BaseModel):
    """Base of the motion-centric models (models/base_model.py:250-303); the input-dict construction for tracking
    evaluation (:255-303) needs the dataset stack and is out of scope."""

    def __init__(self, config, **kwargs):
        super().__init__(config, **kwargs)
        self.save_hyperparameters()
