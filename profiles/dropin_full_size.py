"""MARS.predict drop-in at real sizes with fake PyTorch producers (DINOv2 ViT-L/14-shaped tensors, 24 attention maps)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, marsb200 as mb
dev = torch.device("cuda:0")
g, cdim, regs, heads, layers = 37, 1024, 4, 16, 24
n = g * g
T_tok = 1 + regs + n
for ns, p, hw in ((1, 128, 518), (1, 256, 1024), (5, 256, 518)):
    gen = torch.Generator(device=dev).manual_seed(3)
    feat_s = torch.randn(ns, T_tok, cdim, device=dev, generator=gen)
    feat_q = torch.randn(1, T_tok, cdim, device=dev, generator=gen)
    attn = [torch.softmax(torch.randn(1, heads, T_tok, T_tok, device=dev, generator=gen, dtype=torch.float16).float(), -1).half() for _ in range(layers)]
    masks = mb.synthetic.random_masks(p, hw, hw, gen, dev)
    support = mb.synthetic.random_masks(ns, hw, hw, gen, dev, 0.05, 0.3, dup_frac=0.0)
    vta_raw = torch.rand(33, 33, device=dev, generator=gen)
    img = torch.nn.functional.normalize(torch.randn(p, 768, device=dev, generator=gen), dim=1)
    txt = torch.nn.functional.normalize(torch.randn(768, device=dev, generator=gen), dim=0)

    class FakeDino(torch.nn.Module):
        embed_dim = cdim
        def __init__(self):
            super().__init__(); self.calls = 0
        def forward_features(self, imgs):
            self.calls += 1
            return {"x_prenorm": feat_s if self.calls % 2 == 1 else feat_q}
        def get_last_self_attention(self, img):
            return tuple(attn)
    class FakeText:
        def get_conceptual_information(self, support_images, support_masks): return "thing", ""
    class FakeVTA:
        def compute(self, query_image, fg_label, bg_labels): return vta_raw
    vva_mod = mb.VisualVisualAlignmentModule(FakeDino(), lambda x: x, 14, g, regs, 0.8, 24, dev)
    fm = mb.FilteringMergingModule(None, None, None, alpha=0.85, static_threshold=0.55, dynamic_threshold=0.95, device=dev)
    orig = fm.compute
    fm.compute = lambda **kw: orig(alphaclip_feats=(img, txt), **kw)   # EMD solved on the device (default)
    mars = mb.MARS(FakeText(), FakeVTA(), vva_mod, fm)
    sup_imgs = torch.zeros(1, ns, 3, 8, 8); q_img = torch.zeros(1, 3, 8, 8)
    for it in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pred = mars.predict(sup_imgs, support[None], q_img, masks)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        mars.clear()
    print(f"ns={ns} P={p} {hw}x{hw}: MARS.predict (device EMD, 24 fp16 attention maps) {1e3*(t1-t0):.1f} ms; merged {tuple(pred.shape)} sum {float(pred.sum()):.0f}; selected {int(fm.last['summary'][0,1])}")
