"""The layout-refinement loop body of the reference (``testing/test_render_refine.py``): multi-scale semantic + depth loss
on the 70-channel render (:332-352), the ``PSP_pool_new`` pyramids (:192-215), ``softargmax`` (:20-25) and the gradient
hooks ``fix_grad`` / ``quad_grad`` (:220-230) — plus ``scene_refine``, a decoder-free driver of that loop (BASELINE.json
configs[2]: the layout parameters themselves are optimised; the reference optimises the VAE latent that decodes to them).

The rasterizer is the hot kernel (neural_renderer.render_scene_classes through mesh_render_func); the loss is a handful of
small torch ops on [1,70,256,256] tensors, exactly the ops the reference uses.
"""
import torch
import torch.nn.functional as F

from .diff_render import mesh_render_func

PSP_SIZES = (32, 48, 64, 96)


def softargmax(input_vec, sum_dim, beta=2.0):
    """reference :20-25"""
    idx_vector = torch.cumsum(torch.ones_like(input_vec), dim=sum_dim)
    soft_idx = F.softmax(input_vec * beta, dim=sum_dim)
    return torch.sum(soft_idx * idx_vector, dim=sum_dim) - 1.0


def psp_pool(feats, sizes=PSP_SIZES, output_list=False):
    """PSP_pool_new.forward (:209-215): bilinear(align_corners=True) to each size, then bilinear (align_corners=False, the
    F.upsample default) up to sizes[-1]; concatenated on the channel axis or returned as a list."""
    top = sizes[-1]
    priors = [F.interpolate(F.interpolate(feats, size=(s, s), mode='bilinear', align_corners=True), size=(top, top), mode='bilinear',
                            align_corners=False) for s in sizes]
    return priors if output_list else torch.cat(priors, 1)


def fix_grad(grad_val):
    """:220-225 — average the min-corner and max-corner gradients: boxes translate, sizes stay."""
    g = grad_val.clone().detach()
    avg = g[:, 3:] / 2.0 + g[:, :3] / 2.0
    g[:, 3:] = avg
    g[:, :3] = avg
    return g


def quad_grad(grad_val):
    """:227-230"""
    return grad_val.clone().detach() * 4.0


def refine_targets(target_image):
    """What the reference computes once from the target render (:336-346): pooled depth planes and the per-scale label maps."""
    with torch.no_grad():
        depth = psp_pool(target_image[:, 41:])
        labels = []
        for pooled in psp_pool(target_image[:, 1:41], output_list=True):
            flat = torch.argmax(pooled, dim=1, keepdim=True)
            flat[torch.sum(pooled, dim=1, keepdim=True) < 0.5] = -100
            labels.append(flat[:, 0].long())
    return depth, labels


def refine_loss(iter_image, target_depth, target_labels, size_loss=None):
    """:332-352 — 100 * L1(depth pyramids) * 0.5 + 100 * sum_scales CE(label pyramids) / 800 (+ 2 * size_loss)."""
    iter_image = iter_image.clone()
    null = torch.sum(iter_image[:, 41:], dim=1) < 0.5                      # fill in null regions (:333)
    last = iter_image[:, -1]
    iter_image[:, -1] = torch.where(null, torch.ones_like(last), last)
    depth_loss = F.l1_loss(psp_pool(iter_image[:, 41:]), target_depth) * 0.5
    semantic_loss = 0.0
    for pooled, tgt in zip(psp_pool(iter_image[:, 1:41], output_list=True), target_labels):
        semantic_loss = semantic_loss + F.cross_entropy(pooled, tgt) / 800.0
    loss = depth_loss * 100 + semantic_loss * 100
    if size_loss is not None:
        loss = loss + size_loss * 2.0
    return loss


def scene_refine(boxes, angles, objs, target_boxes=None, target_angles=None, n_iters=200, lr=2e-4, optimizer="adam", callback=None):
    """Refine the layout of ONE scene by gradient descent through the differentiable renderer.

    boxes [n+1, 6] (objects normalised to the room, last row = room box), angles [n+1] (0..24, float), objs [n+1] class ids,
    all on the CUDA device.  The target image is the render of (target_boxes, target_angles) (default: the initial layout
    shifted — callers normally pass the ground-truth layout).  Returns (boxes, angles, losses list).
    Reference loop: testing/test_render_refine.py:279-359 (there the optimised variable is the VAE latent z and the optimiser
    is a re-created SGD; BASELINE.json asks for Adam over the layout)."""
    if boxes.device.type != "cuda":
        raise RuntimeError("scene_refine runs on CUDA only (no CPU fallback)")
    objs_l = [int(o) for o in objs]
    tb = boxes if target_boxes is None else target_boxes
    ta = angles if target_angles is None else target_angles
    with torch.no_grad():
        target, model_ids, sizes, _ = mesh_render_func([tb[i] for i in range(tb.size(0))], [ta[i] for i in range(ta.size(0))], objs_l)
    t_depth, t_labels = refine_targets(target)
    b = boxes.detach().clone().requires_grad_(True)
    a = angles.detach().clone().float().requires_grad_(True)
    opt = torch.optim.Adam([b, a], lr=lr) if optimizer == "adam" else torch.optim.SGD([b, a], lr=lr, nesterov=True, momentum=0.1)
    losses = []
    room = boxes[-1].detach()
    for k in range(n_iters):
        bb = torch.cat([b[:-1], room[None]], 0)                 # boxes_pred[-1] = boxes_gt[-1] (:291)
        bb.register_hook(fix_grad)
        aa = torch.cat([a[:-1], angles[-1:].detach().float()], 0)
        aa.register_hook(quad_grad)
        image, _, _, size_loss = mesh_render_func([bb[i] for i in range(bb.size(0))], [aa[i] for i in range(aa.size(0))], objs_l, model_ids, sizes)
        loss = refine_loss(image, t_depth, t_labels, size_loss)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        losses.append(loss.detach())
        if callback is not None:
            callback(k, loss, image)
    return b.detach(), a.detach(), losses
