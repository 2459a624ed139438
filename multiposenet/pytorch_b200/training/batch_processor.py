"""Compat `batch_processor` (reference: training/batch_processor.py:10-60).

The reference file passes `async=False` to `Tensor.cuda()` (:20-25): `async` is a reserved word since Python 3.7, so the module
is a SyntaxError at import and with it every script that imports it (evaluate/multipose_*_val.py:6, training/multipose_*_train.py).
This mirror keeps the function's contract -- `batch_processor(state, batch) -> (inputs, gts, saved_for_eval)` with
`inputs = [[input_var, subnet_name]]` and `gts = [subnet_name, ...]`, exactly what `Trainer._train_one_epoch` /
`Tester.val` unpack (trainer.py:239-247, tester.py:529-532) -- and spells the keyword `non_blocking`.  Host-side plumbing only:
it moves a batch to `state.params.gpus[0]`; install_dropin() registers it as `training.batch_processor`.
"""
import torch


def batch_processor(state, batch):
    gpus = state.params.gpus
    subnet_name = state.params.subnet_name  # 'detection_subnet' / 'keypoint_subnet' / 'prn_subnet'
    dev = torch.device("cuda", gpus[0])
    grad_ctx = torch.enable_grad() if state.model.training else torch.no_grad()   # :17-19: inference moves under no_grad
    with grad_ctx:
        if subnet_name == "keypoint_subnet":
            inp, heat_temp, heat_weight = batch
            input_var = inp.to(dev)
            heat_weight_var = heat_weight.to(dev, non_blocking=False)
            heat_temp_var = heat_temp.to(dev, non_blocking=False)
            gts = [subnet_name, heat_temp_var, heat_weight_var]
        elif subnet_name == "detection_subnet":
            inp, anno = batch  # anno: [x1, y1, x2, y2, category_id]
            input_var = inp.to(dev)
            gts = [subnet_name, anno.to(dev)]
        else:  # 'prn_subnet'
            inp, label = batch
            input_var = inp.to(dev).float()
            gts = [subnet_name, label.to(dev).float()]
    inputs = [[input_var, subnet_name]]
    saved_for_eval = []
    return inputs, gts, saved_for_eval
