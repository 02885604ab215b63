"""Index arithmetic of the transcribed NLP: where every variable, constraint row and Jacobian value lives.

Pure host logic (no GPU), the Python twin of the layout code in ``csrc/mpx_plan.cu``.  Follows
/root/reference/mpopt/mpopt.py:537-543 (variables), :458 and :617-621 (row order), :189-195 (node ownership)
and :4015-4039 (staircase D).  Used for sharding (which slices of g / values a range of segments owns) and
for slicing trajectories out of z.
"""
from __future__ import annotations

import numpy as np


class PhaseLayout:
    pass


class Layout:
    def __init__(self, program, poly_orders, has_DU, has_mU, has_dU, n_links=0, adaptive=None):
        """``program``: mpopt_b200.program.Program; ``has_*``: per-phase booleans.

        ``adaptive``: None, or ``dict(sw_u=[..], sw_x=[..], mid_residuals=bool)`` for the NLP of ``mpopt_adaptive``
        (mpopt.py:2877-3375): K width variables appended to every phase of z, rows ``[F, C, DU, TC, SW]`` per phase
        (the value offsets ``v*`` then refer to the entries of the base kernels only)."""
        self.po = [int(v) for v in poly_orders]
        self.K = K = len(self.po)
        self.nx, self.nu, self.na, self.P = program.nx, program.nu, program.na, program.n_phases
        nx, nu, na = self.nx, self.nu, self.na
        self.seg_start = np.concatenate([[0], np.cumsum(self.po)]).astype(np.int64)
        self.N = N = int(self.seg_start[-1]) + 1
        self.adaptive = adaptive
        self.nvar = N * (nx + nu) + 2 + na + (K if adaptive else 0)
        self.n_z, self.n_p = self.nvar * self.P, (0 if adaptive else K * self.P)
        po = np.asarray(self.po, dtype=np.int64)
        dcost = po * (po + 1)
        dcost0 = dcost.copy()
        dcost0[0] = (po[0] + 1) ** 2
        # D / mid-point / slope-continuity nonzeros before each segment (index K = totals)
        self.dpre = np.concatenate([[0], np.cumsum(dcost0)])
        self.ipre = np.concatenate([[0], np.cumsum(dcost)])
        self.spre = np.concatenate([[0], np.cumsum(po[:-1] + po[1:] + 1), [0]])
        self.spre[-1] = self.spre[-2]
        self.nnzD, self.nnzI, self.nnzS = int(self.dpre[-1]), int(self.ipre[-1]), int(self.spre[-1])
        self.phases = []
        row = val = 0
        for ph, pp in enumerate(program.phases):
            L = PhaseLayout()
            L.zoff = ph * self.nvar
            L.nc, L.ntc = pp.nc, pp.ntc
            L.has_DU, L.has_mU, L.has_dU = bool(has_DU[ph]), bool(has_mU[ph]), bool(has_dU[ph]) and K > 1
            L.f_next = [len(pre) + len(post) for pre, post in (pp.f_row_layout(s) for s in range(nx))]
            L.c_len = [len(pp.c_row_layout(q)) for q in range(pp.nc)]
            L.tc_len = [len(pp.tc_row_layout(r)) for r in range(pp.ntc)]
            L.gF = row; row += nx * N
            L.gC = row; row += L.nc * N
            L.gDU = row; row += nu * N if L.has_DU else 0
            L.gmU = row; row += nu * (N - 1) if L.has_mU else 0
            L.gdU = row; row += nu * (K - 1) if L.has_dU else 0
            L.gTC = row; row += L.ntc
            if adaptive:  # mpopt.py:3034-3136
                L.gSW = row
                row += 1 + (nu * (N - 1) if adaptive["sw_u"][ph] else 0) + (nx * (N - 1) if adaptive["sw_x"][ph] else 0)
                row += nx * (N - 1) if adaptive["mid_residuals"] else 0
            L.vF, L.vC = [], []
            for s in range(nx):
                L.vF.append(val); val += self.nnzD + N * L.f_next[s]
            for q in range(L.nc):
                L.vC.append(val); val += N * L.c_len[q]
            L.vDU = val; val += nu * self.nnzD if L.has_DU else 0
            L.vmU = val; val += nu * self.nnzI if L.has_mU else 0
            L.vdU = val; val += nu * self.nnzS if L.has_dU else 0
            L.vTC = val; val += sum(L.tc_len)
            self.phases.append(L)
        self.g_events, self.v_events = row, val
        self.n_links = n_links
        self.n_g = row + n_links * (nx + nu + 1)
        self.nnz_full = val + 2 * n_links * (nx + nu + 1)

    # ---- column helpers (mpopt.py:537-543)
    def colX(self, ph, i, s):
        return ph * self.nvar + s * self.N + i

    def colU(self, ph, i, c):
        return ph * self.nvar + (self.nx + c) * self.N + i

    def colT0(self, ph):
        return ph * self.nvar + (self.nx + self.nu) * self.N

    def colTF(self, ph):
        return self.colT0(ph) + 1

    def colA(self, ph, m):
        return self.colT0(ph) + 2 + m

    def colW(self, ph, k):
        """Width variable of segment k (adaptive NLP only, mpopt.py:2938-2945)."""
        return self.colT0(ph) + 2 + self.na + k

    # ---- sharding
    def owned_nodes(self, kb, ke):
        """Nodes whose rows belong to segments [kb, ke): a shared boundary node belongs to the earlier segment."""
        return (0 if kb == 0 else int(self.seg_start[kb]) + 1), int(self.seg_start[ke]) + 1

    def node0_counts(self, kind):
        """Sizes of the pieces of each run of ``shard_runs(kind, 0, ke)`` that belong to global node 0 alone
        (the one row per block that segment 0 owns in addition to its d rows)."""
        d0 = self.po[0]
        out = []
        for L in self.phases:
            if kind == 0:
                out += [1] * (self.nx + L.nc) + ([1] * self.nu if L.has_DU else [])
                out += ([0] * self.nu if L.has_mU else []) + ([0] * self.nu if L.has_dU else [])
            else:
                out += [d0 + 1 + L.f_next[s] for s in range(self.nx)] + [L.c_len[q] for q in range(L.nc)]
                out += ([d0 + 1] * self.nu if L.has_DU else [])
                out += ([0] * self.nu if L.has_mU else []) + ([0] * self.nu if L.has_dU else [])
        return out

    def shard_runs(self, kind, kb, ke):
        """Contiguous (offset, count) runs of g (kind 0), of the unfolded Jacobian values (kind 1) or of grad_f's
        node entries (kind 2) written by the shard that evaluates segments [kb, ke)."""
        N, K, nx, nu = self.N, self.K, self.nx, self.nu
        nb, ne = self.owned_nodes(kb, ke)
        mb, me = int(self.seg_start[kb]), int(self.seg_start[ke])
        db, de = int(self.dpre[kb]), int(self.dpre[ke])
        ib, ie = int(self.ipre[kb]), int(self.ipre[ke])
        ub, ue = kb, min(ke, K - 1)
        sb, se = int(self.spre[ub]), int(self.spre[max(ue, ub)])
        tail = ke == K
        runs = []

        def push(off, cnt):
            if cnt > 0:
                runs.append((int(off), int(cnt)))

        for L in self.phases:
            if kind == 0:
                for s in range(nx):
                    push(L.gF + s * N + nb, ne - nb)
                for q in range(L.nc):
                    push(L.gC + q * N + nb, ne - nb)
                if L.has_DU:
                    for c in range(nu):
                        push(L.gDU + c * N + nb, ne - nb)
                if L.has_mU:
                    for c in range(nu):
                        push(L.gmU + c * (N - 1) + mb, me - mb)
                if L.has_dU:
                    for c in range(nu):
                        push(L.gdU + c * (K - 1) + ub, ue - ub)
                if tail:
                    push(L.gTC, L.ntc)
            elif kind == 1:
                for s in range(nx):
                    push(L.vF[s] + nb * L.f_next[s] + db, (ne - nb) * L.f_next[s] + de - db)
                for q in range(L.nc):
                    push(L.vC[q] + nb * L.c_len[q], (ne - nb) * L.c_len[q])
                if L.has_DU:
                    for c in range(nu):
                        push(L.vDU + c * self.nnzD + db, de - db)
                if L.has_mU:
                    for c in range(nu):
                        push(L.vmU + c * self.nnzI + ib, ie - ib)
                if L.has_dU:
                    for c in range(nu):
                        push(L.vdU + c * self.nnzS + sb, se - sb)
                if tail:
                    push(L.vTC, sum(L.tc_len))
            else:
                for v in range(nx + nu):
                    push(L.zoff + v * N + nb, ne - nb)
        if tail and kind == 0:
            push(self.g_events, self.n_g - self.g_events)
        if tail and kind == 1:
            push(self.v_events, self.nnz_full - self.v_events)
        return runs
