import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_models_gpu as T
for name, B, TT, size in [("early_fusion_mobilenet", 3, 7, 44), ("early_fusion_resnet", 2, 5, 44), ("video_resnet_lstm", 2, 5, 44), ("audio_resnet", 4, 1, 44)]:
    try:
        ref, ours, C = T._case(name)
        wav, mel, lips, labels = T._data(B, size, TT, C)
        ref_in, our_in = T._inputs_for(name, mel, lips)
        ref.train(); ours.train()
        logits_ref = ref(*ref_in)
        loss_ref = torch.nn.functional.cross_entropy(logits_ref, labels)
        loss_ref.backward()
        ours.configure_optimizer()
        loss, logits = ours.train_step(*our_in, labels.cuda(), use_graph=False)
        torch.cuda.synchronize()
        print(f"== {name}: logits err {T._rel(logits, logits_ref):.2e} loss {loss.item():.6f} vs {loss_ref.item():.6f}  launches {ours.launches_per_step()}")
        flat = ours._flat
        rows = []
        for (n, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
            rows.append((T._grad_err(flat.g(p), q.grad, 3e-3), n, q.grad.abs().max().item()))
        nbad = sum(1 for r in rows if r[0] > 3e-3)
        print(f"   {nbad} / {len(rows)} params above 3e-3")
        for e, n, m in rows:
            if e > 1e-3:
                print(f"   {e:.3e}  {n}  (max|g| {m:.2e})")
    except Exception as ex:
        import traceback; traceback.print_exc()
