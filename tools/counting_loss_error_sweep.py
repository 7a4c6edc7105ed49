#!/usr/bin/env python
"""How far the DMPNN / RGIN 'large' loss is from the float64 oracle under each arithmetic variant, over several batches:
    python tools/counting_loss_error_sweep.py            (spawns one process per variant: the switches are read at import)
Variants: library GEMMs / tensor-core GEMM (dn4gl_gemm_f32) x library `Linear, act, Linear` / tensor-core stages."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def child():
    import torch
    from dummynode4graphlearning_b200 import synth, transforms as T
    from dummynode4graphlearning_b200.graph import BatchedGraph
    from dummynode4graphlearning_b200.subgraph_isomorphism.models import DMPNN, RGIN
    from oracle import models as OM
    from helpers import oracle_cfg, rel_err
    import test_models_gpu as TM
    dev = torch.device("cuda:0")
    out = {}
    for name, over in (("DMPNN", dict(node_pred=True, edge_pred=False)), ("RGIN", {})):
        errs = []
        for seed in (5, 6, 7, 8):
            p, g, counts = synth.counting_batch("large", 8, seed=seed)
            cfg = dict(synth.counting_config("large"), add_dummy=True)
            mc = T.process_model_config(cfg)
            pd_ = T.sub_add_dummy(T.to_device(p, dev), cfg["max_npv"], cfg["max_npvl"], cfg["max_npe"], cfg["max_npel"])
            gd_ = T.sub_add_dummy(T.to_device(g, dev), cfg["max_ngv"], cfg["max_ngvl"], cfg["max_nge"], cfg["max_ngel"])
            kw = dict({k: v for k, v in mc.items() if k.startswith("max_")}, hid_dim=64, rep_num_graph_layers=3,
                      rep_num_pattern_layers=3, rep_act_func="leaky_relu", pred_act_func="leaky_relu", pred_net="SumPredictNet",
                      pred_hid_dim=64, emb_net="Equivariant", enc_net="Multihot", filter_net="ScalarFilter", pred_with_enc=True,
                      pred_with_deg=True, rep_rgin_regularizer="bdd", rep_rgin_num_bases=4, pred_return_weights="node",
                      init_neigenv=4.0, init_eeigenv=4.0)
            kw.update(over)
            torch.manual_seed(1)
            model = {"RGIN": RGIN, "DMPNN": DMPNN}[name](**kw)
            with torch.no_grad():
                for n, q in model.named_parameters():
                    if "pred_fc2" in n or "weight_fc2" in n:
                        q.normal_(0.0, 0.1)
            sd = {k: v.clone() for k, v in model.state_dict().items()}
            model = model.to(dev).train()
            pattern, graph = BatchedGraph.from_batch(pd_, dev), BatchedGraph.from_batch(gd_, dev)
            with torch.no_grad():
                loss = TM._loss(model(pattern, graph), torch.from_numpy(counts).to(dev))
                host = lambda b: {k: (v.cpu().numpy() if isinstance(v, torch.Tensor) else v) for k, v in b.items()}
                sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
                l64 = OM.counting_loss(OM.counting_model(sd64, host(pd_), host(gd_), oracle_cfg(name, kw)), torch.from_numpy(counts), rep_reg_w=1e-3)
                l32 = OM.counting_loss(OM.counting_model(sd, host(pd_), host(gd_), oracle_cfg(name, kw)), torch.from_numpy(counts), rep_reg_w=1e-3)
            errs.append((rel_err(loss, l64), rel_err(l32, l64)))
        out[name] = {"gpu_vs_fp64": ["%.1e" % e[0] for e in errs], "fp32_cpu_oracle_vs_fp64": ["%.1e" % e[1] for e in errs]}
    print(json.dumps(out))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for gemm in ("1", "0"):
            for mlp in ("1", "0"):
                env = dict(os.environ, DN4GL_GEMM_TC=gemm, DN4GL_COUNTING_GEMM_TC=gemm, DN4GL_MLP2_TC=mlp)
                r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
                line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                print("gemm_tc=%s mlp2_tc=%s %s" % (gemm, mlp, line[-1] if line else "FAILED: " + r.stderr[-400:]), flush=True)
