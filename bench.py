#!/usr/bin/env python
"""bench.py — headline benchmark of the mmdti_b200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Default workload `hotpath` = BASELINE.json's metric "train molecules/sec (fwd+bwd+contrastive)" on the configs[1]
geometry: the Uni-Mol conformer encoder (15 layers, 64 heads, 512-d, per-GPU batch 128 molecules x 64 atoms, L = 66
tokens, bf16, dropout 0.1) chained the way MM_Model.forward chains the WHOLE hot path of SURVEY.md §8(a)
(models/mm_model.py:545-591): encoder -> InfoNCE against the second modality (a resident random (B, 64, 512) tensor
standing in for the out-of-scope ChemBERTa output) -> cross-modal fusion (CrossAttentionModel, both directions) + masked
mean pooling of both token sequences (--no-fusion: pooling of the encoder output alone) -> FDS.smooth (epoch 1, populated statistics)
-> regression head -> ConR, loss = MSE + 0.1 InfoNCE + 0.1 ConR (tasks/trainer.py:68-69,193), backward, Adam step;
synthetic molecules, random-init weights.  The step is captured once in a CUDA graph and replayed; at N > 1 the
contrastive operands are all-gathered (global-batch negatives, exchange 1) and the gradients all-reduced (exchange 2)
inside the graph.  Metric: train molecules/s, whole job (weak scaling: 128 molecules per GPU).

--workload encoder   configs[1] literally: the encoder alone, loss = sum(all_repr * g)
--workload config3   configs[2]: classification, InfoNCE + SupCon (CT_Single), GLOBAL batch 4096 split 4096/W over the
                     W GPUs (strong scaling), all-gathered negatives
--workload config4   configs[3]: 256 atoms (L = 258), SMILES 256, ConR + FDS, 32 molecules per GPU
--workload config5   configs[4]: contrastive-loss microbench, N = 1K..64K x 512-d vs the tensor roofline (its own metric)

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_PER_GPU, N_ATOMS, LAYERS, HEADS, DIM = 128, 64, 15, 64, 512
L = N_ATOMS + 2
METRIC, UNIT = "train_molecules_per_sec", "molecules/s"
S_SMILES, FDS_BUCKETS = 64, 30
_NAMES = {"encoder": "unimol_encoder_fwd_bwd_15L_64H_512d_b128x64atoms",
          "hotpath": "unimol_encoder_15L_64H_512d_b128x64atoms+infonce+cross_fusion+fds_smooth+conr_fwd_bwd",
          "config3": "unimol_encoder_15L_64H_512d_b128x64atoms+infonce+cross_fusion+supcon_classification_fwd_bwd_global_batch_4096",
          "config4": "unimol_encoder_15L_64H_512d_b128x64atoms+infonce+cross_fusion+fds_smooth+conr_fwd_bwd",
          "config5": "contrastive_similarity_sweep_infonce_supcon_conr_N1K-64K_x512d"}
WORKLOADS = dict(_NAMES)
# per-workload defaults: (molecules per GPU | None = global batch / world, atoms, SMILES length, task, scaling)
SPECS = {"encoder": (128, 64, 64, None, "weak"), "hotpath": (128, 64, 64, "regression", "weak"),
         "config3": (None, 64, 64, "classification", "strong"), "config4": (32, 256, 256, "regression", "weak"),
         "config5": (128, 64, 64, None, "weak")}
GLOBAL_BATCH_CONFIG3 = 4096
CHEMBERTA = False      # --chemberta: the second modality comes from a ChemBERTa-sized RoBERTa encoder in the step
BERT_SHAPE = dict(vocab_size=600, hidden_size=512, num_hidden_layers=6, num_attention_heads=8, intermediate_size=2048,
                  max_position_embeddings=515, pad_token_id=1)       # 512-d stand-in: the reference's heads expect 512 (mm_model.py:493)
FUSION = True          # --no-fusion: pool the encoder output directly (the round-1 step, without SURVEY.md §8 row f2)


def set_shape(batch, n_atoms, smiles_len):
    """Shape of the per-GPU batch (defaults per workload in SPECS; BASELINE configs[1] = 128 x 64 atoms)."""
    global B_PER_GPU, N_ATOMS, L, S_SMILES
    B_PER_GPU, N_ATOMS, S_SMILES = batch, n_atoms, smiles_len
    L = N_ATOMS + 2
    for k in WORKLOADS:
        WORKLOADS[k] = _NAMES[k].replace("b128x64atoms", "b%dx%datoms" % (batch, n_atoms))
        if not FUSION:
            WORKLOADS[k] = WORKLOADS[k].replace("+cross_fusion", "")
        if CHEMBERTA and "+infonce" in WORKLOADS[k] and "chemberta" not in WORKLOADS[k]:
            WORKLOADS[k] = WORKLOADS[k].replace("+infonce", "+chemberta_6L_512d+infonce")


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            p = json.load(fh)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_batch(seed):
    from mmdti_b200.data import synthetic_molecules
    tokens, dist, et, coord = synthetic_molecules(B_PER_GPU, N_ATOMS, seed=seed)
    g = torch.randn(B_PER_GPU, L, DIM, generator=torch.Generator().manual_seed(seed + 7)) * 0.05
    return tokens, dist, et, g, coord


def make_head_batch(seed, n=None, task="regression"):
    """Inputs of the contrastive head: second-modality activations (B, S, 512), targets (standard-scaled regression
    targets, or Bernoulli(0.3) class labels for the classification task), sample weights (mean 1), and FDS statistics of
    a previous epoch."""
    n = B_PER_GPU if n is None else n
    gen = torch.Generator().manual_seed(seed + 11)
    smiles = torch.randn(n, S_SMILES, DIM, generator=gen) * 0.5
    y = torch.randn(n, 1, generator=gen)
    if task == "classification":
        y = (torch.rand(n, 1, generator=gen) < 0.3).float()
    w = torch.rand(n, generator=gen) + 0.5
    w = w / w.mean()
    sg = torch.Generator().manual_seed(99)                      # the same statistics on every rank
    stats = {"running_mean_last_epoch": torch.randn(FDS_BUCKETS, DIM, generator=sg) * 0.1,
             "running_var_last_epoch": torch.rand(FDS_BUCKETS, DIM, generator=sg) + 0.5,
             "smoothed_mean_last_epoch": torch.randn(FDS_BUCKETS, DIM, generator=sg) * 0.1,
             "smoothed_var_last_epoch": torch.rand(FDS_BUCKETS, DIM, generator=sg) + 0.5}
    return smiles, y, w, stats


def smiles_ids(seed, n=None):
    """Token ids of the second modality for --chemberta: random ids under the ragged mask of smiles_mask, padding id 1."""
    n = B_PER_GPU if n is None else n
    gen = torch.Generator().manual_seed(seed + 19)
    ids = torch.randint(4, BERT_SHAPE["vocab_size"], (n, S_SMILES), generator=gen)
    ids[~smiles_mask(seed, n)] = BERT_SHAPE["pad_token_id"]
    return ids


def smiles_mask(seed, n=None):
    """Attention mask of the second modality: ragged lengths in [S/2, S] (tokenizer padding to the longest of the batch)."""
    n = B_PER_GPU if n is None else n
    gen = torch.Generator().manual_seed(seed + 17)
    lens = torch.randint(max(1, S_SMILES // 2), S_SMILES + 1, (n,), generator=gen)
    lens[0] = S_SMILES
    return torch.arange(S_SMILES)[None, :] < lens[:, None]


FDS_CFG = dict(min_value=-3.0, bin_width=0.2, bucket_num=FDS_BUCKETS, bucket_start=0, start_smooth=1)


# ------------------------------------------------------------------ CPU arm
def _reference_modules():
    """The reference's OWN modules (models/mm_model.py, models/transformers.py, models/infonce.py, models/contrastive.py)
    when a copy of the reference tree travelled to this machine (baseline/_ref, made by scripts/install_reference.sh;
    git-ignored) or /root/reference exists; None otherwise (then the arm runs oracle/restate.py, kind "port")."""
    try:
        from oracle import ref_loader
        return ref_loader.load() if ref_loader.available() else None
    except Exception as exc:                                     # a broken copy must not take the bench down
        print("[bench] reference tree not usable (%s); using the oracle port" % exc, file=sys.stderr)
        return None


def cpu_reference_run(steps, warmup, sample_b=32, workload="hotpath"):
    """The reference's CPU implementation of the same path, fp32, all host threads, on a bounded sample of the workload:
    `sample_b` molecules per step, same L / depth / width / loss chain.  kind "reference": the reference's own classes
    (GaussianLayer, NonLinearHead, TransformerEncoderWithPair, InfoNCE, CT_Regress / CT_Single) chained as
    models/mm_model.py:545-591 chains them, Uni-Core layer from oracle/shims; kind "port": oracle/restate.py.
    FDS.smooth always comes from the port (the reference's FDS is CUDA-only, models/fds.py:84).
    Returns (molecules/s, seconds per step, cores, kind)."""
    import torch.nn.functional as F
    from oracle import restate
    from oracle.detw import det_state_dict
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tests_util import slice_shapes
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    task = SPECS[workload][3]
    sd = det_state_dict(slice_shapes(HEADS, DIM, 2048, LAYERS), seed=5)
    tokens, dist, et, g, _ = make_batch(1234)
    sample_b = min(sample_b, B_PER_GPU)
    tokens, dist, et, g = tokens[:sample_b], dist[:sample_b], et[:sample_b], g[:sample_b]
    ref = _reference_modules()
    kind = "reference" if ref is not None else "port"
    torch.manual_seed(5)
    if ref is not None:
        mm, tr = ref["mm_model"], ref["transformers"]
        enc_mods = torch.nn.ModuleDict({
            "embed_tokens": torch.nn.Embedding(31, DIM, 0), "gbf": mm.GaussianLayer(128, 961),
            "gbf_proj": mm.NonLinearHead(128, HEADS, "gelu"),
            "encoder": tr.TransformerEncoderWithPair(encoder_layers=LAYERS, embed_dim=DIM, ffn_embed_dim=2048, attention_heads=HEADS,
                                                     emb_dropout=0.1, dropout=0.1, attention_dropout=0.1, activation_dropout=0.0,
                                                     max_seq_len=512, activation_fn="gelu", no_final_head_layer_norm=True)})
        enc_mods.load_state_dict(sd)
        enc_mods.train()
        enc_params = list(enc_mods.parameters())

        def encode():
            pm = tokens.eq(0)
            x = enc_mods["embed_tokens"](tokens)
            bias = enc_mods["gbf_proj"](enc_mods["gbf"](dist, et)).permute(0, 3, 1, 2).contiguous()
            bias = bias.view(-1, bias.size(-2), bias.size(-1))
            return enc_mods["encoder"](x, padding_mask=pm if pm.any() else None, attn_mask=bias)[0]
    else:
        p = {k: v.requires_grad_(True) for k, v in sd.items()}
        enc_params = list(p.values())

        def encode():
            return restate.unimol_encoder(tokens, dist, et, p, heads=HEADS, n_layers=LAYERS)

    head_params = []
    if task is not None:
        smiles, y, w, stats = (x[:sample_b] if torch.is_tensor(x) else x for x in make_head_batch(1234, task=task))
        head = torch.nn.Linear(DIM, 1)
        if ref is not None:
            inf = ref["infonce"].InfoNCE(DIM, DIM)
            ct = ref["contrastive"]
        else:
            inf = torch.nn.ModuleDict({
                "info_proj_query": torch.nn.Sequential(torch.nn.Linear(DIM, DIM), torch.nn.GELU(), torch.nn.Linear(DIM, 50)),
                "info_proj_positive": torch.nn.Sequential(torch.nn.Linear(DIM, DIM), torch.nn.GELU(), torch.nn.Linear(DIM, 50))})
            pi = {"infonce." + k: v for k, v in inf.named_parameters()}
        head_params = list(inf.parameters()) + list(head.parameters())
        mask = tokens.ne(0).float().unsqueeze(-1)
        if CHEMBERTA:
            # models/mm_model.py:475,562: the reference's own second-modality encoder = Hugging Face RobertaModel
            from transformers import RobertaConfig, RobertaModel
            bert = RobertaModel(RobertaConfig(**BERT_SHAPE)).train()
            ids_c, am_c = smiles_ids(1234, sample_b), smiles_mask(1234, sample_b).long()
            head_params += list(bert.parameters())
        if FUSION:
            # cross-modal fusion + pooling, models/mm_model.py:571-576 (the reference's CrossAttentionModel, else the port)
            img_mask, txt_mask = tokens.ne(0), smiles_mask(1234, sample_b)
            if ref is not None:
                cross = ref["mm_model"].CrossAttentionModel(ref["mm_model"].crossmodal_config(), num_layers=1).train()
                head_params += list(cross.parameters())
            else:
                from tests_util import cross_layer_shapes
                shp = {}
                for side in ("text_attention", "graph_attention"):
                    shp.update(cross_layer_shapes(DIM, 2048, side + ".layer.0."))
                pc = {k: v.requires_grad_(True) for k, v in det_state_dict(shp, seed=6).items()}
                head_params += list(pc.values())

            def pool(rep):
                if ref is not None:
                    a, c = cross(rep, smiles, img_mask, txt_mask)
                else:
                    a, c = restate.cross_modal(rep, smiles, img_mask, txt_mask, pc)
                return restate.fuse_pool(a, c, img_mask, txt_mask)
        else:
            def pool(rep):
                return (rep * mask).sum(1) / mask.sum(1)
    opt = torch.optim.Adam(enc_params + head_params, lr=1e-4, eps=1e-6)   # tasks/trainer.py:160
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        rep = encode()
        if task is None:
            (rep * g).sum().backward()
        else:
            if CHEMBERTA:
                smiles = bert(ids_c, am_c, return_dict=True)[0]
            l_inf = inf(rep, smiles) if ref is not None else restate.infonce_head(rep, smiles, pi)
            pooled = pool(rep)
            if task == "regression":
                feats = restate.fds_smooth(pooled * 1.0, y, 1, stats, FDS_CFG)
                logits = head(feats)
                l_ct = (ct.CT_Regress(feats, y, logits, weights=w, w=0.2) if ref is not None
                        else restate.ct_regress(feats, y, logits, weights=w, w=0.2))
                l_task = F.mse_loss(logits, y)
            else:
                logits = head(pooled)
                l_ct = ct.CT_Single(pooled, y, logits) if ref is not None else restate.ct_single(pooled, y)
                l_task = F.binary_cross_entropy_with_logits(logits, y)
            (l_task + 0.1 * l_inf + 0.1 * l_ct).backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    return sample_b / sec, sec, cores, kind


def base_config(workload, world):
    """The `config` object both arms print (the reference arm adds its sample size)."""
    return {"workload": WORKLOADS[workload], "per_gpu_batch": B_PER_GPU, "global_batch": B_PER_GPU * world,
            "n_atoms": N_ATOMS, "seq_len": L, "smiles_len": S_SMILES, "layers": LAYERS}


def cpu_sample_text(kind, sample_b, sec):
    what = ("the reference's own modules (copy of the reference tree in baseline/_ref, Uni-Core layer from oracle/shims)"
            if kind == "reference" else "oracle/restate.py (no reference tree on this machine)")
    return ("%d of the %d molecules per step (same L=%d, %d layers, same loss chain, fp32, dropout as the reference "
            "configures it), %.1f s/step, %s" % (min(sample_b, B_PER_GPU), B_PER_GPU, L, LAYERS, sec, what))


def run_reference(args, rank, world):
    if rank != 0:
        return
    if args.workload == "config5":
        return run_config5_reference(args)
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    val, sec, cores, kind = cpu_reference_run(steps, warm, args.cpu_sample, args.workload)
    cfg = base_config(args.workload, world)
    cfg["cpu_sample_molecules_per_step"] = min(args.cpu_sample, B_PER_GPU)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": SPECS[args.workload][4], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": cpu_sample_text(kind, args.cpu_sample, sec)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm
def run_ours(args, rank, local_rank, world):
    task, scaling = SPECS[args.workload][3], SPECS[args.workload][4]
    import mmdti_b200
    from mmdti_b200 import _lib, ops
    from mmdti_b200.models.encoder import UnimolEncoder
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist_on = world > 1
    json_out = sys.stdout
    if dist_on:
        import torch.distributed as dist
        # keep stdout to the ONE JSON line: NCCL / torch print banners on fd 1; point fd 1 at stderr and keep a
        # private handle on the real stdout for the result line
        sys.stdout.flush()
        json_out = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    mmdti_b200.set_precision(act="bf16", pair=os.environ.get("MMDTI_PAIR", "bf16"))
    ops.set_seed_rank(rank)           # data-parallel ranks draw different dropout masks (same weights: seed 0 below)
    torch.manual_seed(0)
    model = UnimolEncoder().to(dev).train()
    step_model = model
    use_graph = not args.no_graph
    if world == 1 and B_PER_GPU * L > 150000 and use_graph:
        # config 3 on ONE GPU: 4096 molecules x 66 tokens keep ~120 GB of activations alive for the backward; a captured
        # graph holds them in its private pool on top of the warm-up's, which does not fit in 180 GB.  At this size every
        # kernel runs for hundreds of microseconds, so launching from the host costs nothing.
        use_graph = False
        print("[bench] per-GPU batch %d: running without CUDA-graph capture (activation memory)" % B_PER_GPU, file=sys.stderr)
    # gradients travel as bf16 by default (MMDTI_GRAD_COMM=fp32 keeps fp32 buckets)
    COMM_DTYPE = None if os.environ.get("MMDTI_GRAD_COMM", "bf16") == "fp32" else torch.bfloat16
    if dist_on and not use_graph:
        from torch.nn.parallel import DistributedDataParallel as DDP
        step_model = DDP(model, device_ids=[local_rank], gradient_as_bucket_view=True, bucket_cap_mb=64)
    elif dist_on:
        # replicas start identical (same seed); gradients are exchanged by mmdti_b200.dist.OverlappedGradReducer: bucketed
        # NCCL all-reduces on a communication stream, launched from gradient hooks while the backward is still running
        from mmdti_b200.dist import OverlappedGradReducer
        for prm in model.parameters():
            dist.broadcast(prm.data, src=0)
        # the embedding and K1 (pair-bias) gradients arrive after layer 0's: their own small bucket at the end, so that
        # the bucket holding the first layers is reduced under K1's backward; flat buckets feed FusedAdam directly
        late = [model.embed_tokens.weight] + list(model.gbf.parameters()) + list(model.gbf_proj.parameters()) \
            + list(model.encoder.emb_layer_norm.parameters())
        if task is None:
            reducer = OverlappedGradReducer(model.parameters(), average=True, tail_params=late, keep_flat=not args.torch_adam,
                                            bucket_bytes=int(os.environ.get("MMDTI_BUCKET_MB", "32")) << 20,
                                            comm_dtype=COMM_DTYPE)

    hot = task is not None
    extra_params = []
    if hot:
        import numpy as np
        import torch.nn.functional as F
        from mmdti_b200.models.contrastive import CT_Regress, CT_Single
        from mmdti_b200.models.fds import FDS
        from mmdti_b200.models.infonce import InfoNCE
        torch.manual_seed(5)
        inf = InfoNCE(DIM, DIM).to(dev).train()
        head = torch.nn.Linear(DIM, 1).to(dev)
        smiles, y_h, w_h, stats = make_head_batch(1234 + rank, task=task)
        fds = FDS(feature_dim=DIM, raw_data=np.array([0.0, 1.0]), col_data=None, using_scale=False, bucket_num=FDS_BUCKETS).to(dev)
        fds.min_value, fds.bin_width = FDS_CFG["min_value"], FDS_CFG["bin_width"]
        for k, v in stats.items():
            getattr(fds, k).copy_(v)
        d_smiles = smiles.to(dev)
        extra_params = list(inf.parameters()) + list(head.parameters())
        d_txt_mask = smiles_mask(1234 + rank).to(dev)
        if CHEMBERTA:
            from transformers import RobertaConfig
            from mmdti_b200.models.encoder import ChembertaEncoder
            bert = ChembertaEncoder(RobertaConfig(**BERT_SHAPE)).to(dev).train()
            ids_h = smiles_ids(1234 + rank)
            extra_params += list(bert.parameters())
        if FUSION:
            from mmdti_b200.models.cross_modal import CrossAttentionModel, crossmodal_config
            cross = CrossAttentionModel(crossmodal_config(), num_layers=1).to(dev).train()
            extra_params += list(cross.parameters())
        dp_ctx = None
        if dist_on:
            from mmdti_b200.dist import DataParallelCtx
            for prm in extra_params:
                dist.broadcast(prm.data, src=0)
            # exchange 1: every rank scores its rows against the all-gathered global batch; the gradient exchange averages
            dp_ctx = DataParallelCtx()
            inf.dp = fds.dp = dp_ctx
        if dist_on and use_graph:
            reducer = OverlappedGradReducer(list(model.parameters()) + extra_params, average=True, tail_params=late,
                                            keep_flat=not args.torch_adam,
                                            bucket_bytes=int(os.environ.get("MMDTI_BUCKET_MB", "32")) << 20,
                                            comm_dtype=COMM_DTYPE)
        elif dist_on:
            raise SystemExit("--workload hotpath at N > 1 needs the graphed step (drop --no-graph)")

    tokens, dmat, et, g, coord = make_batch(1234 + rank)
    # --inputs pair (default): the reference's batch format (src_tokens, src_distance, src_edge_type);
    # --inputs coords: tokens + coordinates only, the pair features are computed on the device
    host_inputs = (tokens, dmat, et) if args.inputs == "pair" else (tokens, coord)
    n_enc_in = len(host_inputs)
    if hot:
        host_inputs = host_inputs + (y_h, w_h)             # targets and sample weights travel with the batch
        if CHEMBERTA:
            host_inputs = host_inputs + (ids_h,)           # and the SMILES token ids
    pin = [t.pin_memory() for t in host_inputs]
    dev_inputs = [t.to(dev) for t in host_inputs]
    d_g = g.to(dev)
    # Adam(eps 1e-6) as in tasks/trainer.py:160: mmdti_b200.optim.FusedAdam = one launch per step that also refreshes
    # the bf16 shadows of the encoder's GEMM weights (--torch-adam: torch.optim.Adam(fused=True) + per-step cast pass)
    if args.torch_adam:
        opt = torch.optim.Adam(list(model.parameters()) + extra_params, lr=1e-4, eps=1e-6, fused=True, capturable=use_graph)
    else:
        from mmdti_b200.optim import FusedAdam
        flat = dist_on and use_graph
        opt = FusedAdam(list(model.parameters()) + extra_params, lr=1e-4, eps=1e-6, shadows=model.encoder.use_external_lowp(),
                        grad_scale=reducer.grad_scale if flat else 1.0, grad_source=reducer.reduced_grad if flat else None)

    params = [prm for prm in model.parameters() if prm.requires_grad]

    def barrier():
        if dist_on:
            dist.barrier()
        torch.cuda.synchronize()

    def full_step(*inp):
        rep = step_model(*inp[:n_enc_in]) if args.inputs == "pair" else step_model(inp[0], src_coord=inp[1])
        if hot:
            y_d, w_d = inp[n_enc_in], inp[n_enc_in + 1]
            rep = rep.float()
            out_bert = bert(inp[n_enc_in + 2], d_txt_mask) if CHEMBERTA else d_smiles      # f4 (mm_model.py:562)
            l_inf = inf(rep, out_bert)                                     # a7: InfoNCE against the second modality
            if FUSION:                                                      # f2: cross-modal fusion + masked pooling (mm_model.py:571-576)
                img_mask = inp[0].ne(0)
                pooled = cross.forward_pooled(rep, out_bert, img_mask, d_txt_mask)
            else:
                mk = inp[0].ne(0).unsqueeze(-1).float()
                pooled = (rep * mk).sum(1) / mk.sum(1)
            if task == "regression":
                feats = fds.smooth(pooled * 1.0, y_d, 1)                    # a11: in place on the pooled features
                logits = head(feats)
                l_ct = CT_Regress(feats, y_d, logits, weights=w_d, w=0.2, dp=dp_ctx)     # a8: ConR sees the smoothed features
                loss = F.mse_loss(logits, y_d) + 0.1 * l_inf + 0.1 * l_ct   # tasks/trainer.py:68-69,193
            else:                                                           # classification: no FDS (mm_model.py:580)
                logits = head(pooled)
                l_ct = CT_Single(pooled, y_d, logits, dp=dp_ctx)            # a9: SupCon over the global batch
                loss = F.binary_cross_entropy_with_logits(logits, y_d) + 0.1 * l_inf + 0.1 * l_ct
        else:
            loss = (rep * d_g).sum()
        loss.backward()
        if dist_on and use_graph:
            reducer.finish()                # exchange 2: join the overlapped NCCL all-reduces (captured in the graph)
        opt.step()
        return loss.detach()

    def step_eager():
        loss = full_step(*dev_inputs)
        opt.zero_grad(set_to_none=True)
        return loss

    if use_graph:
        from mmdti_b200.graph import GraphedStep
        opt.zero_grad(set_to_none=True)
        graphed = GraphedStep(full_step, dev_inputs, device=dev, params=list(model.parameters()) + extra_params,
                              capture_error_mode="thread_local" if dist_on else "global")

        def step_resident():
            return graphed(*dev_inputs)

        primed = []

        def step_e2e():
            # input pipeline of a training loop: every step copies ITS batch from pinned host memory (H2D inside the timed
            # region, one copy per step), but the copy of step i+1 is issued on a copy stream while step i replays;
            # then the replay and the D2H read of the loss
            if not primed:
                graphed.prefetch(*pin)
                primed.append(True)
            out = graphed()
            graphed.prefetch(*pin)
            return float(out.item())
    else:
        step_resident = step_eager

        def step_e2e():
            loss = full_step(*(x.to(dev, non_blocking=True) for x in pin))
            opt.zero_grad(set_to_none=True)
            return float(loss.item())                  # device -> host read of the step's result

    def timed(fn, steps, timeline=False):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if timeline:
            _lib.start_timeline()
        n0 = _lib.launch_count
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        t = a.elapsed_time(b) * 1e-3
        tl = _lib.stop_timeline() if timeline else None
        launches = _lib.launch_count - n0
        if dist_on:
            tt = torch.tensor([t], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            t = float(tt.item())
        return t, launches, tl

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_res, launches, tl = timed(step_resident, args.steps, timeline=not use_graph)
    clocks = sampler.stop() if rank == 0 else None
    for _ in range(2):
        step_e2e()
    t_e2e, _, _ = timed(step_e2e, args.steps)
    tl_steps = args.steps
    if use_graph:
        # per-kernel CUDA-event times cannot be taken inside a graph replay: one extra EAGER pass of the same
        # step (same kernels, same arguments), used only for the roofline / breakdown fields
        graphed.close()
        opt.zero_grad(set_to_none=True)
        tl_steps = min(args.steps, 3)
        # every kernel is timed ALONE on the launching stream: the weight-gradient GEMMs, which the step itself runs on a
        # side stream underneath the backward chain, are issued in-stream for this pass, so that a kernel's CUDA-event
        # bracket does not contain the SM time another stream's kernel took from it
        ops.overlap_wgrad = False
        for _ in range(2):
            step_eager()
        _, _, tl = timed(step_eager, tl_steps, timeline=True)
        ops.overlap_wgrad = True

    if rank == 0:
        mols = B_PER_GPU * world * args.steps
        peak, peak_src = peaks()
        # dominant kernel of ours by live CUDA-event time inside the timed region
        esz = 2 if mmdti_b200.config.pair_dtype() != torch.float32 else 4
        Lp = ops.pair_ld(L)
        nel = B_PER_GPU * HEADS * L * Lp
        qkvo = B_PER_GPU * L * DIM * 2
        alg = {"mmdti_pair_attn_fwd": 2 * nel * esz + 4 * qkvo,            # read P, write P', q,k,v in, o out
               "mmdti_pair_attn_bwd": 3 * nel * esz + 9 * qkvo}            # read S, dP'; write dP; q,k,v,o,dO in; dq,dk,dv out
        breakdown = {k: {"calls_per_step": n / tl_steps, "ms_per_step": 1e3 * s / tl_steps} for k, (n, s) in (tl or {}).items()}
        dom = max((k for k in breakdown if k in alg), key=lambda k: breakdown[k]["ms_per_step"], default=None)
        roof = None
        if dom:
            n, s = tl[dom]
            # the last layer's backward reads no dP' (MM-DTI discards the pair output): 1 of 15 launches
            bytes_per_launch = alg[dom] - (nel * esz / LAYERS if dom.endswith("bwd") else 0)
            ach = bytes_per_launch / (s / n) / 1e9
            traffic = None
            try:        # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
                with open(os.path.join(ROOT, "profiles", "k2_dram_traffic.json")) as fh:
                    tr = json.load(fh)[dom]
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
            except Exception:
                pass
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": traffic, "traffic_source": "profiles/k2_dram_traffic.json (ncu --set full, bytes per launch)",
                    "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                    "avg_launch_us": 1e6 * s / n, "share_of_step": (s / tl_steps) / (t_res / args.steps),
                    "timing": "CUDA events around each launch on the launching stream"
                              + (" (separate eager pass with every kernel in-stream: the timed region replays a CUDA graph whose "
                                 "weight-gradient GEMMs run on a side stream)" if use_graph else "")}
        h2d = sum(t.numel() * t.element_size() for t in pin)
        line = {
            "metric": METRIC, "value": mols / t_res, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {**base_config(args.workload, world),
                       "pair_dtype": os.environ.get("MMDTI_PAIR", "bf16"), "dropout": 0.1,
                       "optimizer": "Adam(eps=1e-6), " + ("torch fused" if args.torch_adam else "mmdti FusedAdam (one launch, writes bf16 weight shadows)"),
                       "cuda_graph": bool(use_graph),
                       "inputs": ("src_tokens + src_distance + src_edge_type (reference batch format)" if args.inputs == "pair"
                                  else "src_tokens + src_coord (pair features computed on the device, mmdti_featurise)"),
                       "parallelism": "dp%d" % world,
                       "grad_exchange": (None if world == 1 else (("bucketed NCCL all-reduce (32 MB of parameters per bucket, %s on the wire) overlapped with the backward, inside the graph"
                                                                   % os.environ.get("MMDTI_GRAD_COMM", "bf16")) if use_graph
                                                                  else "DistributedDataParallel (NCCL)")),
                       "l2": "no flush: the per-step working set (15 x %.0f MB pair tensors + activations) exceeds the 126 MB L2"
                             % (nel * esz / 1e6)},
            "e2e": {"value": mols / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                    "ms_per_step": 1e3 * t_e2e / args.steps,
                    "pipeline": ("H2D of step i+1 overlaps the replay of step i (GraphedStep.prefetch), loss read back every step"
                                 if use_graph else "H2D, step, loss read back; no overlap")},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
            "kernel_breakdown": breakdown,
        }
        if world == 1 and not args.no_cpu_baseline:
            val, sec, cores, kind = cpu_reference_run(steps=3, warmup=1, sample_b=args.cpu_sample, workload=args.workload)
            line["cpu_baseline"] = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": cpu_sample_text(kind, args.cpu_sample, sec) + ", 3 timed steps"}
        print(json.dumps(line), file=json_out, flush=True)
    if dist_on:
        # Tear down in a fixed order: drain the device, make sure every rank is done, destroy the captured graph (it
        # holds NCCL kernels and keeps the communicator's resources referenced), remove the gradient hooks, and only
        # then destroy the process group.  A watchdog bounds the teardown: if the communicator destruction does not
        # return within 30 s the process reports it on stderr and leaves with the result already printed.
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        if use_graph:
            graphed.graph.reset()
            del graphed
            reducer.remove()
        torch.cuda.synchronize()

        def _stuck():
            print("[bench rank %d] destroy_process_group did not return within 30 s; exiting" % rank, file=sys.stderr, flush=True)
            os._exit(0)

        wd = threading.Timer(30.0, _stuck)
        wd.daemon = True
        wd.start()
        dist.destroy_process_group()
        wd.cancel()


# ------------------------------------------------------------------ config 5: contrastive-loss microbench
def _config5_cases(N, D, dev, gen):
    f = torch.randn(N, D, device=dev, generator=gen)
    f2 = torch.randn(N, D, device=dev, generator=gen)
    y = torch.randn(N, 1, device=dev, generator=gen)
    yhat = y + 0.3 * torch.randn(N, 1, device=dev, generator=gen)
    cls = torch.randint(0, 10, (N, 1), device=dev, generator=gen)
    return f, f2, y, yhat, cls


def run_config5(args):
    """BASELINE configs[4]: InfoNCE / SupCon / ConR on N x 512-d embeddings, N = 1K..64K, forward and forward + backward,
    against the measured dense bf16 peak.  One JSON line; `value` = InfoNCE forward + backward TFLOP/s at the largest N
    (flops = 12 N^2 D: two directions x (similarity + recompute + two gradient GEMMs))."""
    import mmdti_b200
    from mmdti_b200 import _lib
    from mmdti_b200.models import contrastive as ctm
    from mmdti_b200.models import infonce as infm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback for the product path)"
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    mmdti_b200.set_precision(act="bf16")
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peak, peak_src = float(json.load(fh)["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst: each case is timed alone)"
    except Exception:
        peak, peak_src = 1590.0, "fallback (B200_PROFILING.md 1.59 PFLOP/s)"
    D = 512
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def timeit(fn, iters):
        for _ in range(max(args.warmup, 3)):
            fn()
        torch.cuda.synchronize()
        tot = 0.0
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b) * 1e-3
        return tot / iters

    sampler = ClockSampler(0)
    sampler.start()
    rows, N = [], 1024
    n0 = _lib.launch_count
    while N <= args.nmax:
        gen = torch.Generator(device=dev).manual_seed(N)
        f, f2, y, yhat, cls = _config5_cases(N, D, dev, gen)
        iters = max(2, min(args.steps, int(4e12 / (N * N * D)) + 2))
        cases = {"infonce": (lambda a: infm.info_nce(a, f2), 4, 8),          # two directions
                 "supcon": (lambda a: ctm.CT_Single(a, cls, None), 2, 4),
                 "conr": (lambda a: ctm.CT_Regress(a, y, yhat), 2, 4)}
        for name, (fn, ffl, bfl) in cases.items():
            a = f.clone().requires_grad_(True)

            def fwd():
                with torch.no_grad():
                    fn(a)

            def fwdbwd():
                a.grad = None
                fn(a).backward()

            tf, tb = timeit(fwd, iters), timeit(fwdbwd, iters)
            flf, flb = ffl * N * N * D, (ffl + bfl) * N * N * D
            rows.append({"loss": name, "N": N, "D": D, "fwd_ms": tf * 1e3, "fwd_tflops": flf / tf / 1e12, "fwd_frac": flf / tf / 1e12 / peak,
                         "fwdbwd_ms": tb * 1e3, "fwdbwd_tflops": flb / tb / 1e12, "fwdbwd_frac": flb / tb / 1e12 / peak})
            print("[config5] %-8s N=%6d fwd %8.3f ms %7.1f TF/s (%.2f)  fwd+bwd %8.3f ms %7.1f TF/s (%.2f)"
                  % (name, N, tf * 1e3, flf / tf / 1e12, flf / tf / 1e12 / peak, tb * 1e3, flb / tb / 1e12, flb / tb / 1e12 / peak),
                  file=sys.stderr, flush=True)
        N *= 2
    clocks = sampler.stop()
    top = [r for r in rows if r["loss"] == "infonce"][-1]
    line = {"metric": "contrastive_loss_fwd_bwd_tflops", "value": top["fwdbwd_tflops"], "unit": "TFLOP/s", "n_gpus": 1,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": top["fwdbwd_ms"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOADS["config5"], "N_headline": top["N"], "D": D, "loss_headline": "infonce",
                       "l2": "flushed (256 MB write) before every timed call"},
            "roofline": {"kernel": "sim_tc_kernel (tcgen05 similarity engine, phase 1 + backward)", "bound": "tensor",
                         "achieved": top["fwdbwd_tflops"], "peak": peak, "unit": "TFLOP/s", "frac": top["fwdbwd_frac"],
                         "traffic": None, "peak_source": peak_src},
            "gpu_launches": _lib.launch_count - n0, "clocks": clocks, "sweep": rows}
    if not args.no_cpu_baseline:
        line["cpu_baseline"] = _config5_cpu()
    print(json.dumps(line), flush=True)


def _config5_cpu(n=2048, d=512):
    """CPU arm of config 5: the reference's InfoNCE arithmetic (models/infonce.py:70-98 restated in oracle/restate.py, or
    the reference's own function when its tree is present) forward + backward on N = 2048 x 512-d, all host threads."""
    from oracle import restate
    ref = _reference_modules()
    fn = ref["infonce"].info_nce if ref is not None else (lambda q, k: restate.info_nce(q, k, 0.1))
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(1)
    q = torch.randn(n, d, generator=g).requires_grad_(True)
    k = torch.randn(n, d, generator=g).requires_grad_(True)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        fn(q, k).backward()
        ts.append(time.perf_counter() - t0)
    sec = min(ts[1:])
    return {"value": 12.0 * n * n * d / sec / 1e12, "unit": "TFLOP/s", "cores": cores, "kind": "reference" if ref is not None else "port",
            "sample": "InfoNCE forward + backward at N = %d x %d-d (%.3f s), fp32" % (n, d, sec)}


def run_config5_reference(args):
    cb = _config5_cpu()
    line = {"impl": "reference", "metric": "contrastive_loss_fwd_bwd_tflops", "value": cb["value"], "unit": "TFLOP/s", "n_gpus": args.gpus,
            "steps": 3, "warmup": 1, "ms_per_step": None, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": WORKLOADS["config5"], "N_headline": 2048, "D": 512, "loss_headline": "infonce"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="use torch.optim.Adam(fused=True) instead of mmdti_b200.optim.FusedAdam")
    ap.add_argument("--inputs", default="pair", choices=["pair", "coords"],
                    help="host batch format of the e2e path: the reference's (tokens, distance, edge_type) or (tokens, coordinates)")
    ap.add_argument("--workload", default="hotpath", choices=sorted(WORKLOADS),
                    help="hotpath (default) = encoder + InfoNCE + FDS.smooth + ConR in the step on the configs[1] geometry; "
                         "encoder = configs[1] literally; config3 / config4 / config5 = BASELINE configs[2..4]")
    ap.add_argument("--batch", type=int, default=None, help="molecules per GPU (default: per workload, 128 for configs[1])")
    ap.add_argument("--n-atoms", type=int, default=None, help="atoms per molecule; L = n_atoms + 2 <= 264")
    ap.add_argument("--smiles-len", type=int, default=None, help="length of the second-modality sequence")
    ap.add_argument("--cpu-sample", type=int, default=32, help="molecules per step of the CPU arm's bounded sample")
    ap.add_argument("--nmax", type=int, default=65536, help="config5: largest N of the sweep")
    ap.add_argument("--chemberta", action="store_true", help="run a ChemBERTa-sized RoBERTa encoder (6 layers, 512-d) on SMILES token ids "
                    "inside the step instead of the resident stand-in tensor (SURVEY.md 8 row f4)")
    ap.add_argument("--no-fusion", action="store_true", help="leave the cross-modal fusion block (SURVEY.md 8 row f2) out of the step")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    n_ranks = max(world, args.gpus)
    global FUSION, CHEMBERTA
    FUSION = not args.no_fusion
    CHEMBERTA = args.chemberta
    b0, a0, s0 = SPECS[args.workload][:3]
    if b0 is None:                                      # config 3: the GLOBAL batch is fixed (strong scaling)
        if GLOBAL_BATCH_CONFIG3 % n_ranks:
            raise SystemExit("config3: the global batch %d is not divisible by %d GPUs" % (GLOBAL_BATCH_CONFIG3, n_ranks))
        b0 = GLOBAL_BATCH_CONFIG3 // n_ranks
    set_shape(args.batch or b0, args.n_atoms or a0, args.smiles_len or s0)
    if args.impl == "reference":
        run_reference(args, rank, n_ranks)
        return
    if args.workload == "config5":
        if rank == 0:
            run_config5(args)
        return
    if world == 1 and args.gpus > 1:
        # not launched under torchrun: re-exec one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
