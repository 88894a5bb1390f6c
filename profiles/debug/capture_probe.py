import os, sys, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from types import SimpleNamespace as NS
import imp_b200
from imp_b200 import survival, step as S
from imp_b200.registry import build_model
import imp_b200.umeml_gan
dev = torch.device("cuda", 0)
B, N, P = 4, 2048, 32
cfg = NS(DATASET=NS(ROOT=".", PATH=NS(DIM=512), OMIC=NS(DIM=3354)),
         MODEL=NS(DROPOUT=0.25, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small",
                  UMEML=NS(PROTOTYPES=P, REGISTERS=3, GENE_GROUP_INDEXES=None, IMPORTANCE_LOG="defer")), TRAINER=NS(PREC="fp32"))
os.chdir("/tmp")
model = build_model("umeml_gan", verbose=False, cfg=cfg, num_classes=4, omic_sizes=1000).to(dev).train()
x = torch.randn(B * N, 512, device=dev).bfloat16()
cu = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=dev)
batch = {"x_packed": x, "cu_seqlens": cu, "max_len": N, "omic": torch.rand(B, 3354, device=dev), "patient_id": None}
y = torch.randint(0, 4, (B,), device=dev); c = torch.randint(0, 2, (B,), device=dev)
tok = torch.randn(B, 7, 256, device=dev, requires_grad=True)
tokp = torch.randn(B, 33, 256, device=dev, requires_grad=True)
from imp_b200 import token_tail as TT
def pinv_only():
    a = torch.softmax(tok @ tok.transpose(1, 2), dim=-1)
    return TT.iterative_pinv(a).sum()
def nystrom_noconv():
    m = model.path_decoder.attn
    old = m.residual; m.residual = False
    try:
        return m(tok).sum()
    finally:
        m.residual = old
stages = {
  "hot": lambda: (imp_b200.model.IMPHotPath.forward(model, batch)["p_proto"]).sum(),
  "pinv": pinv_only,
  "nystrom_noconv": nystrom_noconv,
  "translayer": lambda: model.path_decoder(tok).sum(),
  "bottleattn": lambda: model.bottleattn(tokp, tok)[0].sum(),
  "tail_eval_like": lambda: model._fuse_and_classify(tokp, tok, None).sum(),
  "full": lambda: (lambda out: survival.nll_loss_new(out, y, c) + out[5] + out[1])(model(batch)),
}
for name, fn in stages.items():
    try:
        gs = S.GraphedStep(None).capture_fn(fn, list(model.parameters()), dev)
        gs.replay(); torch.cuda.synchronize()
        print(name, "capture ok", float(gs.loss))
        gs.close()
    except Exception:
        print(name, "capture FAILED")
        print(traceback.format_exc().splitlines()[-12:])
        try:
            torch.cuda.synchronize()
        except Exception:
            pass
