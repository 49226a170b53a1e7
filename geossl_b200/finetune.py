"""Inner steps of the fine-tune callers of the encoder.

``md17_losses`` is the body of /root/reference/examples/finetune_md17.py::train (:30-48): energy from the readout,
force = -dE/dpos by autograd with ``create_graph=True``, and the weighted L1 objective whose backward differentiates
through the force (double backward; served by the closed-under-differentiation cfconv primitives in ops.py).
``lba_loss`` is the body of finetune_lba.py::train (:33-47).
"""
import torch
from torch.autograd import grad


def _readout(args, model, batch_data, positions):
    x = batch_data.x[:, 0] if batch_data.x.dim() == 2 else batch_data.x
    if args.model_3d == "schnet":
        return model(x, positions, batch_data.batch, num_graphs=getattr(batch_data, "n_graphs", None))
    if args.model_3d == "painn":
        return model(x, positions, batch_data.radius_edge_index, batch_data.batch,
                     num_graphs=getattr(batch_data, "n_graphs", None))
    raise Exception("3D model {} not included.".format(args.model_3d))


def md17_losses(args, batch_data, model, graph_pred_linear, criterion, energy_coeff=0.05, force_coeff=0.95):
    """Returns ``(loss, pred_energy, pred_force)``; ``batch_data`` carries ``y`` (B,) and ``force`` (N,3) in extras or
    as attributes.  Coefficients default to submit_finetune_md17_schnet.sh's 0.05 / 0.95."""
    positions = batch_data.positions
    positions.requires_grad_()
    molecule_3D_repr = _readout(args, model, batch_data, positions)
    pred_energy = (graph_pred_linear(molecule_3D_repr) if graph_pred_linear is not None else molecule_3D_repr).squeeze(1)
    pred_force = -grad(outputs=pred_energy, inputs=positions, grad_outputs=torch.ones_like(pred_energy),
                       create_graph=True, retain_graph=True)[0]
    y = getattr(batch_data, "y", None)
    y = batch_data.extras["y"] if y is None else y
    f = getattr(batch_data, "force", None)
    f = batch_data.extras["force"] if f is None else f
    loss = energy_coeff * criterion(pred_energy, y) + force_coeff * criterion(pred_force, f)
    return loss, pred_energy, pred_force


def md17_train_step(args, batch_data, model, graph_pred_linear, criterion, optimizer, **coeffs):
    loss, _, _ = md17_losses(args, batch_data, model, graph_pred_linear, criterion, **coeffs)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()


def lba_loss(args, batch, model, graph_pred_linear, criterion):
    molecule_3D_repr = _readout(args, model, batch, batch.positions)
    pred = (graph_pred_linear(molecule_3D_repr) if graph_pred_linear is not None else molecule_3D_repr).squeeze()
    y = getattr(batch, "y", None)
    return criterion(pred, batch.extras["y"] if y is None else y)


def lba_train_step(args, batch, model, graph_pred_linear, criterion, optimizer):
    loss = lba_loss(args, batch, model, graph_pred_linear, criterion)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return loss.detach()
