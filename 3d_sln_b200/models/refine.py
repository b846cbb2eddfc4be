"""The layout-refinement loop body of the reference (``testing/test_render_refine.py``): multi-scale semantic + depth loss
on the 70-channel render (:332-352), the ``PSP_pool_new`` pyramids (:192-215), ``softargmax`` (:20-25) and the gradient
hooks ``fix_grad`` / ``quad_grad`` (:220-230) — plus ``scene_refine``, a decoder-free driver of that loop (BASELINE.json
configs[2]: the layout parameters themselves are optimised; the reference optimises the VAE latent that decodes to them).

The rasterizer is the hot kernel (neural_renderer.render_scene_classes through mesh_render_func); the loss is a handful of
small torch ops on [1,70,256,256] tensors, exactly the ops the reference uses.
"""
import torch
import torch.nn.functional as F


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


def scene_refine(boxes, angles, objs, target_boxes=None, target_angles=None, n_iters=200, lr=2e-4, use_graph=True, callback=None):
    """Refine the layout of ONE scene by gradient descent through the differentiable renderer (BASELINE.json configs[2]).

    boxes [n+1, 6] (objects normalised to the room, last row = room box), angles [n+1] (0..24, float), objs [n+1] class ids,
    on the CUDA device.  The target image is the render of (target_boxes, target_angles).  Returns (boxes, angles, losses).
    Reference loop: testing/test_render_refine.py:279-359 (there the optimised variable is the VAE latent z that decodes to the
    layout and the optimiser is a re-created SGD; BASELINE.json asks for Adam over the layout itself).  With use_graph the whole
    iteration (render, multi-scale loss, backward, Adam) is one CUDA-graph replay (RefineStep)."""
    if boxes.device.type != "cuda":
        raise RuntimeError("scene_refine runs on CUDA only (no CPU fallback)")
    tb = boxes if target_boxes is None else target_boxes
    ta = angles if target_angles is None else target_angles
    step = RefineStep(boxes, angles, objs, tb, ta, lr=lr, use_graph=use_graph)
    losses = []
    for k in range(n_iters):
        loss = step.step()
        losses.append(loss.detach().clone())
        if callback is not None:
            callback(k, loss, step)
    return step.b.detach().clone(), step.a.detach().clone(), losses


class RefineStep(object):
    """One refinement iteration (render -> multi-scale loss -> backward -> Adam) of a fixed scene as a replayable CUDA graph.

    The layout (boxes [n+1,6], angles [n+1]) lives in static device tensors that the graph updates in place; ``step()`` replays
    the graph and returns the (device) loss of that iteration.  Same arithmetic as ``scene_refine`` — the graph removes the
    ~600 kernel-launch / Python overheads per iteration that otherwise dominate (the rasterizer itself takes < 1 ms)."""

    def __init__(self, boxes, angles, objs, target_boxes, target_angles, lr=2e-4, use_graph=True, library=None):
        from . import diff_render as dr
        dev = boxes.device
        if dev.type != "cuda":
            raise RuntimeError("RefineStep runs on CUDA only (no CPU fallback)")
        lib = library if library is not None else dr.mesh_library(dev)
        self.static = dr.SceneStatic(objs, boxes[-1], lib, dev)
        with torch.no_grad():
            target, tsize = dr.render_static(self.static, target_boxes.to(dev), target_angles.to(dev).float())
        self.t_depth, self.t_labels = refine_targets(target)
        self.size_target = tsize.detach()
        self.room_row = boxes[-1:].detach().clone()
        self.angle_room = angles[-1:].detach().float().clone()
        self.b = boxes.detach().clone().requires_grad_(True)
        self.a = angles.detach().float().clone().requires_grad_(True)
        self.opt = torch.optim.Adam([self.b, self.a], lr=lr, capturable=True)
        self.loss = torch.zeros((), device=dev)
        self.graph = None
        self._dr = dr
        if use_graph:
            self.capture()

    def _iteration(self):
        bb = torch.cat([self.b[:-1], self.room_row], 0)
        bb.register_hook(fix_grad)
        aa = torch.cat([self.a[:-1], self.angle_room], 0)
        aa.register_hook(quad_grad)
        image, size = self._dr.render_static(self.static, bb, aa)
        size_loss = ((size - self.size_target) ** 2).mean(dim=1).sum()        # :98: sum over objects of mse(size, size of the first render)
        loss = refine_loss(image, self.t_depth, self.t_labels, size_loss)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        self.opt.step()
        self.loss.copy_(loss.detach())

    def capture(self):
        s = torch.cuda.Stream(self.b.device)
        s.wait_stream(torch.cuda.current_stream(self.b.device))
        b0, a0 = self.b.detach().clone(), self.a.detach().clone()
        with torch.cuda.stream(s):
            for _ in range(3):
                self._iteration()
        torch.cuda.current_stream(self.b.device).wait_stream(s)
        torch.cuda.synchronize(self.b.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        self.reset(b0, a0)
        return self

    def reset(self, boxes, angles):
        """Restart from a layout (also clears the Adam moments)."""
        with torch.no_grad():
            self.b.copy_(boxes); self.a.copy_(angles.float())
            for st in self.opt.state.values():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.zero_()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._iteration()
        return self.loss
