import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gadm_b200 as G
DEV = "cuda:0"
torch.manual_seed(0)
for k, N in ((32, 200), (64, 300), (128, 400), (256, 800), (300, 1000), (1024, 3000)):
    phi = torch.randn(N, k, device=DEV)
    Kd = phi.double().T @ phi.double() + 0.5 * torch.eye(k, device=DEV, dtype=torch.float64)
    phi_t = G.transpose(phi)
    gram = G.gemm_tn(phi_t, phi_t, lower_only=True, diag_add=0.5)
    torch.cuda.synchronize()
    gl = torch.tril(gram.double()); gerr = float((gl - torch.tril(Kd)).abs().max() / Kd.abs().max())
    sc = G.TrakScorer(0.5).factor_(gram.clone())
    torch.cuda.synchronize()
    L = torch.tril(sc.L.double())
    Lref = torch.linalg.cholesky(Kd)
    rel = float((L @ L.T - Kd).abs().max() / Kd.abs().max())
    lerr = float((L - Lref).abs().max() / Lref.abs().max())
    # where is L wrong? per 128-block max error
    nb = (k + 127) // 128
    blk = [[float((L - Lref)[i*128:(i+1)*128, j*128:(j+1)*128].abs().max()) if j <= i else 0.0 for j in range(nb)] for i in range(nb)]
    rows = torch.randn(70, k, device=DEV)
    z = sc.solve_rows(rows)
    want = torch.linalg.solve(Kd, rows.double().T).T
    serr = float((z.double() - want).abs().max() / want.abs().max())
    # forward-only check using linv blocks
    nblk = (k + 127) // 128
    blocks = sc.blocks.view(torch.float32).view(2, nblk, 128, 128)
    d0 = min(k, 128)
    li = blocks[0, 0, :d0, :d0].double()
    linv_err = float((li @ Lref[:d0, :d0] - torch.eye(d0, device=DEV, dtype=torch.float64)).abs().max())
    lit_err = float((blocks[1, 0, :d0, :d0].double() - li.T).abs().max())
    print(f"k={k} N={N} info={int(sc.info)} gram_err={gerr:.2e} LLt_err={rel:.2e} L_err={lerr:.2e} linv_err={linv_err:.2e} linvT_err={lit_err:.2e} solve_err={serr:.2e} U_err={float((sc.U.double()-sc.L.double().T).abs().max()):.2e}")
    if nb <= 3: print("   block errs", blk)
