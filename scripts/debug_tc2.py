"""Ad-hoc GPU check of the pipelined tcgen05 kernel against the oracle (development aid, not part of the test-suite)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kagnn_b200 as kb
from kagnn_b200 import ops, _lib as L
from oracle import kagnn_oracle as K

def sd_cpu(m): return {k: v.detach().cpu() for k, v in m.state_dict().items()}
def run(m, *a):
    with torch.no_grad():
        return m(*a).cpu()

torch.manual_seed(0)
for variant in (0, 1):
    ops.set_tc_variant(variant)
    print("== variant", variant)
    for (G, k, fin, fout, n) in [(5,3,16,16,128),(5,3,64,64,300),(4,2,16,16,128),(4,2,7,12,60),(4,2,12,12,60),(3,1,16,16,128),(5,3,128,64,1000)]:
        torch.manual_seed(G*100+fin)
        m = kb.KANLinear(fin, fout, grid_size=G, spline_order=k)
        x = torch.randn(n, fin) * 0.9
        ref = K._kan_layer_from_sd(sd_cpu(m), "", x)
        c0 = ops.launch_counters()
        y = run(m.cuda(), x.cuda())
        c1 = ops.launch_counters()
        print(f"KANLinear G{G} k{k} {fin}->{fout} n{n}: err {K.rel_err(y, ref):.2e}  tc2+{c1['tc2']-c0['tc2']} tc+{c1['tc']-c0['tc']}")
    # bases only: zero base weight; and base only: zero spline weight
    for which in ("spline_only", "base_only"):
        for k in (3, 2):
            torch.manual_seed(5)
            m = kb.KANLinear(16, 16, grid_size=4, spline_order=k)
            with torch.no_grad():
                if which == "spline_only": m.base_weight.zero_()
                else: m.spline_weight.zero_()
            x = torch.randn(128, 16) * 0.9
            ref = K._kan_layer_from_sd(sd_cpu(m), "", x)
            y = run(m.cuda(), x.cuda())
            print(f"{which} k{k}: err {K.rel_err(y, ref):.2e}")
    # GCN-style aggregation with SiLU pre-affine then KAN
    torch.manual_seed(1)
    n, e, f = 60, 212, 12
    ei = torch.randint(0, n, (2, e))
    for mode in ("gin", "gcnfuse", "pool"):
        lin = kb.KANLinear(f, 12, grid_size=4, spline_order=2)
        x = torch.randn(n, f)
        sd = sd_cpu(lin)
        if mode == "gin":
            agg_ref = x + torch.zeros(n, f).index_add_(0, ei[1], x[ei[0]])
        if mode == "gin":
            from kagnn_b200.graph import get_graph
            g = get_graph(ei.cuda(), n)
            y = ops.fused_layer(ops.AggSpec(L.AGG_GIN, x.cuda(), g.rowptr, g.col, self_scale=1.0), n, lin.cuda().kernel_specs()).cpu()
            ref = K._kan_layer_from_sd(sd, "", agg_ref)
        elif mode == "gcnfuse":
            from kagnn_b200.graph import get_graph
            g = get_graph(ei.cuda(), n)
            w, sw = g.gcn_weights()
            bias = torch.randn(f)
            y = ops.fused_layer(ops.AggSpec(L.AGG_WEIGHTED, x.cuda(), g.rowptr, g.col, edge_weight=w, self_weight=sw), n,
                                lin.cuda().kernel_specs(), pre=ops.Affine(shift=bias.cuda(), act=L.ACT_SILU)).cpu()
            h = K.gcn_conv(x, ei, lambda t: t, bias) if hasattr(K, "gcn_conv") else None
            ref = K._kan_layer_from_sd(sd, "", torch.nn.functional.silu(h)) if h is not None else y
        else:
            batch = torch.sort(torch.randint(0, 9, (n,)))[0]
            ptr = ops.segment_ptr(batch.cuda(), 9)
            y = ops.fused_layer(ops.AggSpec(L.AGG_SEGMENT_MEAN, x.cuda(), rowptr=ptr), 9, lin.cuda().kernel_specs()).cpu()
            pooled = torch.zeros(9, f).index_add_(0, batch, x) / torch.bincount(batch, minlength=9).clamp(min=1).unsqueeze(1)
            ref = K._kan_layer_from_sd(sd, "", pooled)
        print(f"{mode}: err {K.rel_err(y, ref):.2e}")
ops.set_tc_variant(0)
