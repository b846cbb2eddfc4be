"""The layout-refinement loop body of the reference (``testing/test_render_refine.py``): multi-scale semantic + depth loss
on the 70-channel render (:332-352), the ``PSP_pool_new`` pyramids (:192-215), ``softargmax`` (:20-25) and the gradient
hooks ``fix_grad`` / ``quad_grad`` (:220-230) — plus ``scene_refine``, a decoder-free driver of that loop (BASELINE.json
configs[2]: the layout parameters themselves are optimised; the reference optimises the VAE latent that decodes to them).

The rasterizer is the hot kernel (neural_renderer.render_scene_classes through mesh_render_func).  The loss exists twice:
``refine_loss`` — the reference's own torch ops (≈150 small kernels per iteration, forward + autograd; kept as the checker and for
CPU tensors) — and ``FusedRefineLoss`` — the same arithmetic as six launches of csrc/refine_loss.cu (forward and the gradient
w.r.t. the render in one call), which RefineStep uses.
"""
import ctypes

import torch
import torch.nn.functional as F

from .. import _lib


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


class _SizeLossFn(torch.autograd.Function):
    """weight * sum_j mse(size_j, target_j) (diff_render.py:98 summed over the objects, test_render_refine.py:352) with an analytic
    backward: 5 small launches instead of the 14 of the autograd chain sub -> pow -> mean -> sum -> mul."""

    @staticmethod
    def forward(ctx, size, target, weight):
        diff = size - target
        ctx.save_for_backward(diff)
        ctx.k = 2.0 * weight / size.size(1)
        return (diff * diff).sum() * (weight / size.size(1))

    @staticmethod
    def backward(ctx, g):
        (diff,) = ctx.saved_tensors
        return diff * (g * ctx.k), None, None


class _FusedRefineLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, image, holder):
        lib = _lib.load()
        img = image.contiguous().float()
        need_grad = bool(ctx.needs_input_grad[0])      # False under no_grad / for detached images: forward-only call
        d_image = torch.empty_like(img) if need_grad else None
        loss3 = torch.empty(3, device=img.device, dtype=torch.float32)
        _lib.check(lib.sln_refine_loss(img.data_ptr(), img.size(-1), holder.sizes_c, holder.n_sem, holder.n_dep, holder.t_depth.data_ptr(),
                                       holder.t_labels.data_ptr(), holder.counts_c, loss3.data_ptr(),
                                       d_image.data_ptr() if need_grad else None, holder.ws.data_ptr(), holder.ws.numel(),
                                       _lib.cur_stream(img.device)), "refine_loss")
        ctx.d_image = d_image
        holder.last_terms = loss3
        return loss3[0]

    @staticmethod
    def backward(ctx, g):
        return (ctx.d_image * g if ctx.d_image is not None else None), None


class FusedRefineLoss(object):
    """refine_loss (test_render_refine.py:332-352) for a FIXED target as one library call (csrc/refine_loss.cu).

    target_depth [1, 4*n_dep, top, top], target_labels: 4 x [1, top, top] int64 (refine_targets).  ``loss = f(image, size_loss)``
    with image [1, 1+n_sem+n_dep, S, S] on CUDA; the gradient w.r.t. the image is produced by the same call (no atomics)."""

    def __init__(self, target_depth, target_labels, n_sem=40, sizes=PSP_SIZES):
        dev = target_depth.device
        if dev.type != "cuda":
            raise RuntimeError("FusedRefineLoss runs on CUDA only (refine_loss is the torch restatement)")
        if len(sizes) != 4 or len(target_labels) != 4:
            raise ValueError("FusedRefineLoss: four pyramid levels expected")
        self.lib = _lib.load()
        self.n_sem, self.n_dep = int(n_sem), int(target_depth.size(1)) // 4
        self.sizes = tuple(int(s) for s in sizes)
        self.sizes_c = (ctypes.c_int32 * 4)(*self.sizes)
        self.t_depth = target_depth.detach().contiguous().float()
        self.t_labels = torch.stack([t.reshape(self.sizes[-1], self.sizes[-1]) for t in target_labels]).contiguous().long()
        self.counts_c = (ctypes.c_float * 4)(*[float((t >= 0).sum().item()) for t in self.t_labels])
        self.image_size = None
        self.ws = None
        self.last_terms = None      # device [3]: total, depth term, semantic term of the last call

    def __call__(self, iter_image, size_loss=None):
        if iter_image.device.type != "cuda":
            raise RuntimeError("FusedRefineLoss needs a CUDA image (no CPU fallback)")
        if iter_image.dim() != 4 or iter_image.size(0) != 1 or iter_image.size(1) != 1 + self.n_sem + self.n_dep or iter_image.size(2) != iter_image.size(3):
            raise ValueError("FusedRefineLoss: image must be [1, %d, S, S], got %s" % (1 + self.n_sem + self.n_dep, tuple(iter_image.shape)))
        S = iter_image.size(-1)
        if self.image_size != S:
            nbytes = self.lib.sln_refine_loss_workspace_bytes(S, self.sizes_c, self.n_sem, self.n_dep)
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=iter_image.device)
            self.image_size = S
        loss = _FusedRefineLossFn.apply(iter_image, self)
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

    def __init__(self, boxes, angles, objs, target_boxes, target_angles, lr=2e-4, use_graph=True, library=None, fused_loss=True, fused_scene=True):
        from . import diff_render as dr
        dev = boxes.device
        if dev.type != "cuda":
            raise RuntimeError("RefineStep runs on CUDA only (no CPU fallback)")
        lib = library if library is not None else dr.mesh_library(dev)
        self.static = dr.SceneStatic(objs, boxes[-1], lib, dev)
        with torch.no_grad():
            target, tsize = dr.render_static(self.static, target_boxes.to(dev), target_angles.to(dev).float(), fused=fused_scene)
        self.t_depth, self.t_labels = refine_targets(target)
        self.fused_loss = FusedRefineLoss(self.t_depth, self.t_labels) if fused_loss else None
        self.fused_scene = fused_scene   # False: scene assembly and compositing through the torch-op restatements (checker)
        self.size_target = tsize.detach()
        self.room_row = boxes[-1:].detach().clone()
        self.angle_room = angles[-1:].detach().float().clone()
        # the layout lives in ONE flat leaf (boxes | angles): one gradient buffer, one library Adam launch (sln_adam_step,
        # the same kernel as the VAE train step) instead of the ~15 launches of torch's capturable Adam
        n = boxes.size(0)
        self.flat = torch.cat([boxes.detach().float().reshape(-1), angles.detach().float().reshape(-1)]).clone().requires_grad_(True)
        self.flat.grad = torch.zeros_like(self.flat)
        self.n_rows = n
        self.lr = lr
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        self.step_count = torch.zeros(1, device=dev, dtype=torch.int64)
        self.lib = _lib.load()
        self.loss = torch.zeros((), device=dev)
        self.graph = None
        self._dr = dr
        if use_graph:
            self.capture()

    @property
    def b(self):
        """boxes [n+1, 6]: a view of the flat layout leaf"""
        return self.flat[:6 * self.n_rows].view(self.n_rows, 6)

    @property
    def a(self):
        """angles [n+1]: a view of the flat layout leaf"""
        return self.flat[6 * self.n_rows:]

    def _iteration(self):
        b, a = self.b, self.a
        if self.fused_scene:
            # the room row owns no mesh: its gradient is exactly 0 and Adam leaves it alone, so the layout leaf itself is the render
            # input; fix_grad / quad_grad (:220-230) are applied inside the assembly's backward kernel
            image, size = self._dr.render_static(self.static, b, a, fused=True, refine_hooks=True)
        else:
            bb = torch.cat([b[:-1], self.room_row], 0)
            bb.register_hook(fix_grad)
            aa = torch.cat([a[:-1], self.angle_room], 0)
            aa.register_hook(quad_grad)
            image, size = self._dr.render_static(self.static, bb, aa, fused=False)
        if self.fused_loss is not None:
            loss = self.fused_loss(image) + _SizeLossFn.apply(size, self.size_target, 2.0)
        else:
            size_loss = ((size - self.size_target) ** 2).mean(dim=1).sum()    # :98: sum over objects of mse(size, size of the first render)
            loss = refine_loss(image, self.t_depth, self.t_labels, size_loss)
        self.flat.grad.zero_()
        loss.backward()
        _lib.check(self.lib.sln_adam_step(self.flat.data_ptr(), self.flat.grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), self.flat.numel(),
                                          self.lr, 0.9, 0.999, 1e-8, 0.0, 1.0, self.step_count.data_ptr(), 1, _lib.cur_stream(self.flat.device)),
                   "adam_step")       # torch.optim.Adam defaults (BASELINE configs[2]: Adam lr 2e-4 over boxes + angles)
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
            self.m.zero_(); self.v.zero_(); self.step_count.zero_()

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._iteration()
        return self.loss


def bias_box_head(model, box=(0.3, 0.0, 0.3, 0.6, 0.4, 0.6), weight_scale=0.1):
    """Synthetic stand-in for a trained checkpoint (none is available offline): a random-init decoder predicts boxes around 0 with
    non-positive sizes, which render nothing and give the refinement loop no gradient.  Shrink the box head's last Linear and set its
    bias to a visible box, so that every object renders and d loss / d z is non-trivial.  Benchmarks and tests only."""
    last = [m for m in model.box_net if isinstance(m, torch.nn.Linear)][-1]
    with torch.no_grad():
        last.weight.mul_(weight_scale)
        last.bias.copy_(torch.tensor(box, dtype=last.bias.dtype, device=last.bias.device))
    return model


class ReferenceRefineStep(object):
    """The reference's OWN refinement iteration (testing/test_render_refine.py:279-359), one CUDA-graph replay per iteration:

        z (leaf) -> model.decoder(z, objs, triples, attributes) [eval-mode BatchNorm]            :287
        boxes_pred.register_hook(fix_grad); boxes_pred[-1] = boxes_gt[-1]                         :288-291
        angles = softargmax(angles_pred, 1) + randn(n) / 10; register_hook(quad_grad); [-1] = gt  :293-298
        render (scene assembly, ONE rasterization, compositing)                                   :324
        null-fill, PSP pyramids, 100 * depth + 100 * semantic + 2 * size loss                     :332-352
        backward THROUGH THE DECODER to z and to the model parameters                             :356
        SGD(nesterov, momentum 0.1) re-created every iteration (:286) => stateless: p -= lr * 1.1 * grad,
            lr = 2e-4 for z, learning_rate / 10 for the parameters                                :286,357

    (RefineStep above optimises the layout itself with Adam — BASELINE.json configs[2]; this class is the loop the reference runs.)
    The gradient hooks are applied inside the scene-assembly backward kernel; the size targets are the object sizes of the first
    render of the prediction (:325-328), the target image is the render of the ground-truth layout (:318-321)."""

    def __init__(self, model, z, objs, triples, attributes, boxes_gt, angles_gt, lr_z=2e-4, lr_model=1e-5, noise=True, use_graph=True,
                 library=None, update_model=True):
        from . import diff_render as dr
        dev = z.device
        if dev.type != "cuda":
            raise RuntimeError("ReferenceRefineStep runs on CUDA only (no CPU fallback)")
        if model.training:
            raise RuntimeError("ReferenceRefineStep: the reference refines with model.eval() (test_render_refine.py:264)")
        self._dr, self.model, self.dev = dr, model, dev
        self.objs, self.triples, self.attrs = objs.to(dev), triples.to(dev), attributes.to(dev)
        self.boxes_gt, self.angles_gt = boxes_gt.to(dev).float(), angles_gt.to(dev).float()
        lib = library if library is not None else dr.mesh_library(dev)
        self.static = dr.SceneStatic(objs, self.boxes_gt[-1], lib, dev)
        with torch.no_grad():
            target, _ = dr.render_static(self.static, self.boxes_gt, self.angles_gt, fused=True)          # "Rendering gt" :318-321
        self.t_depth, self.t_labels = refine_targets(target)
        self.fused_loss = FusedRefineLoss(self.t_depth, self.t_labels)
        self.z = z.detach().clone().float().requires_grad_(True)
        self.noise_on = bool(noise)
        self.noise = torch.zeros(objs.size(0), device=dev)
        self.lr_z, self.lr_model, self.update_model = float(lr_z), float(lr_model), bool(update_model)
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.loss = torch.zeros((), device=dev)
        self.boxes_pred = torch.zeros(objs.size(0), 6, device=dev)
        self.angles_pred = torch.zeros(objs.size(0), device=dev)
        with torch.no_grad():                                                                            # size targets = first render of the prediction
            b, a = self._layout(sample_noise=False)
            _, size = dr.render_static(self.static, b, a, fused=True)
        self.size_target = size.detach().clone()
        self.graph = None
        if use_graph:
            self.capture()

    def _layout(self, sample_noise=True):
        boxes_pred, angles_pred = self.model.decoder(self.z, self.objs, self.triples, self.attrs)
        n = boxes_pred.size(0)
        boxes = torch.cat([boxes_pred[:-1], self.boxes_gt[-1:]], 0)                                      # boxes_pred[-1] = boxes_gt[-1]
        ang = softargmax(angles_pred, sum_dim=1)
        if self.noise_on and sample_noise:
            self.noise.normal_()
            ang = ang + self.noise / 10.0
        ang = torch.cat([ang[:-1], self.angles_gt[-1:]], 0)                                              # angles_pred_idx2[-1] = angles[-1]
        return boxes, ang

    def _iteration(self):
        boxes, ang = self._layout()
        image, size = self._dr.render_static(self.static, boxes, ang, fused=True, refine_hooks=True)     # hooks fused into the assembly backward
        loss = self.fused_loss(image) + _SizeLossFn.apply(size, self.size_target, 2.0)
        if self.z.grad is not None:
            self.z.grad.zero_()
        for p in self.params:                                                                            # optimizer.zero_grad() :355
            if p.grad is not None:
                p.grad.zero_()
        loss.backward()
        with torch.no_grad():
            # torch.optim.SGD(nesterov=True, momentum=0.1) with a FRESH momentum buffer: buf = g, update = g + 0.1 * buf
            self.z.add_(self.z.grad, alpha=-self.lr_z * 1.1)
            if self.update_model:
                ps = [p for p in self.params if p.grad is not None]
                if ps:
                    torch._foreach_add_(ps, [p.grad for p in ps], alpha=-self.lr_model * 1.1)
            self.loss.copy_(loss.detach())
            self.boxes_pred.copy_(boxes.detach()); self.angles_pred.copy_(ang.detach())

    def capture(self):
        s = torch.cuda.Stream(self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        z0 = self.z.detach().clone()
        p0 = [p.detach().clone() for p in self.params]
        with torch.cuda.stream(s):
            for _ in range(3):
                self._iteration()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        with torch.no_grad():                                  # undo the warm-up / capture iterations
            self.z.copy_(z0)
            for p, q in zip(self.params, p0):
                p.copy_(q)
        return self

    def step(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._iteration()
        return self.loss
