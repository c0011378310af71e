"""Learning-rate schedule of the reference trainer for the fused fine-tuning path.

trainers/train.py:187 wraps its AdamW in transformers.get_linear_schedule_with_warmup(optimizer, num_warmup_steps,
num_training_steps): linear ramp 0 -> lr over the warm-up steps, then linear decay to 0 at t_total.  The fused path
(OrderingEngine.adamw_step / BertForOrdering.finetune_step) takes the learning rate per call, so the schedule is a pure
function of the optimizer-step index (0-based: the value the scheduler holds while step number `step` is applied)."""


def linear_schedule_with_warmup(base_lr, step, num_warmup_steps, num_training_steps):
    if step < num_warmup_steps:
        return base_lr * float(step) / float(max(1, num_warmup_steps))
    return base_lr * max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))
