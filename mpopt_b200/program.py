"""Trace an OCP's per-phase callables and generate the CUDA node functors.

One ``PhaseProgram`` per phase holds the traced DAGs of
``dynamics / path_constraints / running_costs / terminal_constraints /
terminal_costs`` (reference call sites: /root/reference/mpopt/mpopt.py:186-206 and
:277-298), their symbolic partials and the *structural* pattern of those partials.
``Program.cuda_source()`` emits one ``struct MpxPh<k>`` per phase whose static
members are what the hand-written kernels in ``csrc/mpx_kernels.cuh`` need:
sizes, the per-row layout of the dynamic Jacobian entries, and straight-line
``__device__`` functions evaluating values + packed partials at one node.

Variable ids inside a phase (columns of the pattern matrices handed to the C ABI):
``x_0..x_{nx-1}, u_0..u_{nu-1}, a_0..a_{na-1}`` -> ``0 .. nx+nu+na-1``;
terminal functions use ``xf_0.., x0_0.., tf, t0, a_0..`` -> ``0 .. 2nx+2+na-1``.
"""
from __future__ import annotations

import hashlib

from . import trace as tr
from .ca import Vec


class PhaseProgram:
    def __init__(self, ocp, phase: int):
        nx, nu, na = ocp.nx, ocp.nu, ocp.na
        self.nx, self.nu, self.na, self.phase = nx, nu, na, phase
        tag = f"p{phase}_"
        self.x = [tr.var(f"{tag}x{s}") for s in range(nx)]
        self.u = [tr.var(f"{tag}u{c}") for c in range(nu)]
        self.a = [tr.var(f"{tag}a{m}") for m in range(na)]
        self.t = tr.var(f"{tag}t")
        self.node_vars = self.x + self.u + self.a
        x, u, a = Vec(self.x), Vec(self.u), Vec(self.a)

        # ---- node functions (mpopt.py:186-206)
        f = tr.flatten(ocp.get_dynamics(phase)(x, u, self.t, a))
        if len(f) != nx:
            raise ValueError(f"dynamics of phase {phase} returned {len(f)} components, expected {nx}")
        self.f = [tr.as_expr(e) for e in f]
        self.c = []
        if ocp.has_path_constraints(phase):  # mpopt.py:171
            self.c = [tr.as_expr(e) for e in tr.flatten(ocp.get_path_constraints(phase)(x, u, self.t, a))]
        self.L = tr.as_expr(tr.flatten(ocp.get_running_costs(phase)(x, u, self.t, a))[0])
        self.nc = len(self.c)

        def jac(outs):
            ent, dt = [], []
            for r, e in enumerate(outs):
                g = tr.gradient(e, self.node_vars + [self.t])
                ent += [(r, v, d) for v, d in enumerate(g[:-1]) if not d.is_value(0.0)]
                dt.append(g[-1])
            return ent, dt

        self.jf, self.ft = jac(self.f)  # [(row s, var id, Expr)], [d f_s / d t]
        self.jc, self.ct = jac(self.c)
        gl, lt = jac([self.L])
        self.gL, self.Lt = [(v, d) for _, v, d in gl], lt[0]
        self.f_nz = [not e.is_value(0.0) for e in self.f]
        self.L_nz = not self.L.is_value(0.0)

        # ---- terminal functions (mpopt.py:277-298)
        self.xf = [tr.var(f"{tag}xf{s}") for s in range(nx)]
        self.x0 = [tr.var(f"{tag}xi{s}") for s in range(nx)]
        self.tf, self.t0 = tr.var(f"{tag}tf"), tr.var(f"{tag}t0")
        self.term_vars = self.xf + self.x0 + [self.tf, self.t0] + self.a
        targs = (Vec(self.xf), self.tf, Vec(self.x0), self.t0, a)
        self.tc = []
        if ocp.has_terminal_constraints(phase):  # mpopt.py:284
            self.tc = [tr.as_expr(e) for e in tr.flatten(ocp.get_terminal_constraints(phase)(*targs))]
        self.ntc = len(self.tc)
        self.M = tr.as_expr(tr.flatten(ocp.get_terminal_costs(phase)(*targs))[0])
        self.jtc = []
        for r, e in enumerate(self.tc):
            g = tr.gradient(e, self.term_vars)
            self.jtc += [(r, v, d) for v, d in enumerate(g) if not d.is_value(0.0)]
        self.gM = [(v, d) for v, d in enumerate(tr.gradient(self.M, self.term_vars)) if not d.is_value(0.0)]

        # ---- Hessian of the Lagrangian (CasADi's nlp_hess_l, implicit in mpopt.py:757).  At one node
        #      lag = h (sw L - sum_s lamF_s Sx_s f_s) + sum_q lamC_q c_q  with the segment width h and the multipliers as
        #      extra symbols; second derivatives w.r.t. W = (x.., u.., a.., t, h), lower triangle.  The kernel maps
        #      (t, h) -> (T0, TF) with the (linear) chain rule.
        self.hs, self.sw = tr.var(f"{tag}hs"), tr.var(f"{tag}sw")
        self.lamF = [tr.var(f"{tag}lf{s}") for s in range(nx)]
        self.lamC = [tr.var(f"{tag}lc{q}") for q in range(self.nc)]
        self.sxs = [tr.var(f"{tag}sx{s}") for s in range(nx)]
        phi = tr.mul(self.sw, self.L)
        for s in range(nx):
            phi = tr.sub(phi, tr.mul(tr.mul(self.lamF[s], self.sxs[s]), self.f[s]))
        lag = tr.mul(self.hs, phi)
        for q in range(self.nc):
            lag = tr.add(lag, tr.mul(self.lamC[q], self.c[q]))
        self.hess_vars = self.node_vars + [self.t, self.hs]
        self.hw = self._lower_hessian(lag, self.hess_vars)  # [(a, b, Expr)], b <= a
        # ---- mid-point residual rows of the widths-as-variables NLP (mpopt.py:3084-3136): their Lagrangian term is
        #      w_k sum_m (mu_m . DI X - h_k psi_m) with psi = sum_s mu_s Sx_s f_s evaluated at the interpolated mid
        #      point; value, gradient and Hessian of psi w.r.t. (x.., u.., a..) (the multipliers ride in lamF)
        psi = tr.as_expr(0.0)
        for s in range(nx):
            psi = tr.add(psi, tr.mul(tr.mul(self.lamF[s], self.sxs[s]), self.f[s]))
        self.psi = psi
        self.rg = [(v, d) for v, d in enumerate(tr.gradient(psi, self.node_vars)) if not d.is_value(0.0)]
        self.rh = self._lower_hessian(psi, self.node_vars)
        self.lamT = [tr.var(f"{tag}lt{r}") for r in range(self.ntc)]
        theta = tr.mul(self.sw, self.M)
        for r in range(self.ntc):
            theta = tr.add(theta, tr.mul(self.lamT[r], self.tc[r]))
        self.ht = self._lower_hessian(theta, self.term_vars)

    @staticmethod
    def _lower_hessian(scalar, wrt):
        g = tr.gradient(scalar, wrt)
        ent = []
        for a, ga in enumerate(g):
            if ga.is_value(0.0):
                continue
            for b, d in enumerate(tr.gradient(ga, wrt[: a + 1])):
                if not d.is_value(0.0):
                    ent.append((a, b, d))
        return ent

    def pat_hw(self):
        n = len(self.hess_vars)
        p = [[0] * n for _ in range(n)]
        for a, b, _ in self.hw:
            p[a][b] = 1
        return p

    def pat_hf(self):
        """[nv][nv] lower-triangle pattern of the second derivatives of sum_s mu_s f_s (residual rows, adaptive NLP)."""
        n = self.nv
        p = [[0] * n for _ in range(n)]
        for a, b, _ in self.rh:
            p[a][b] = 1
        return p

    def pat_ht(self):
        n = len(self.term_vars)
        p = [[0] * n for _ in range(n)]
        for a, b, _ in self.ht:
            p[a][b] = 1
        return p

    def hess_layout(self):
        """Categories and slots of the node Hessian entries, shared by the code generator and Layout:
        cat 0 YY (slot = ordinal), 1 AY (ordinal), 2 AA (corner index), 3 tY, 4 hY (slot = ordinal of q among the node
        variables coupled to T0 / TF), 5 tA, 6 hA (slot = m), 7 tt, 8 ht."""
        ny, na, nv = self.nx + self.nu, self.na, self.nv
        it, ih = nv, nv + 1
        ty = sorted({b for a, b, _ in self.hw if a in (it, ih) and b < ny})
        cat, slot = [], []
        n_yy = n_ay = 0
        for a, b, _ in self.hw:
            if a < ny:
                cat.append(0), slot.append(n_yy)
                n_yy += 1
            elif a < nv and b < ny:
                cat.append(1), slot.append(n_ay)
                n_ay += 1
            elif a < nv:
                m, n = a - ny, b - ny
                cat.append(2), slot.append(3 + 2 * na + m * (m + 1) // 2 + n)
            elif b < ny:
                cat.append(3 if a == it else 4), slot.append(ty.index(b))
            elif b < nv:
                cat.append(5 if a == it else 6), slot.append(b - ny)
            elif a == it:
                cat.append(7), slot.append(0)
            else:
                assert b == it, "d2/dh2 is structurally zero"
                cat.append(8), slot.append(0)
        return dict(cat=cat, slot=slot, ty=ty, n_yy=n_yy, n_ay=n_ay, n_corner=3 + 2 * na + na * (na + 1) // 2)

    # ---- structural patterns handed to the C ABI (uint8 row-major)
    @property
    def nv(self):
        return self.nx + self.nu + self.na

    def pat_f(self):
        p = [[0] * self.nv for _ in range(self.nx)]
        for r, v, _ in self.jf:
            p[r][v] = 1
        return p

    def pat_c(self):
        p = [[0] * self.nv for _ in range(self.nc)]
        for r, v, _ in self.jc:
            p[r][v] = 1
        return p

    def pat_tc(self):
        p = [[0] * len(self.term_vars) for _ in range(self.ntc)]
        for r, v, _ in self.jtc:
            p[r][v] = 1
        return p

    def f_t(self):
        return [int(not d.is_value(0.0)) for d in self.ft]

    def c_t(self):
        return [int(not d.is_value(0.0)) for d in self.ct]

    # ---- row layouts (must agree with csrc/mpx_plan.cpp: build_structure)
    def f_row_layout(self, s):
        """Sorted extras of row F(s, .) excluding the D block: list of ('x'|'u'|'T0'|'TF'|'a', index)."""
        nx, nu = self.nx, self.nu
        pat = self.pat_f()[s]
        pre = [("x", sp) for sp in range(s) if pat[sp]]
        post = [("x", sp) for sp in range(s + 1, nx) if pat[sp]]
        post += [("u", c) for c in range(nu) if pat[nx + c]]
        if self.f_nz[s]:
            post += [("T0", 0), ("TF", 0)]
        post += [("a", m) for m in range(self.na) if pat[nx + nu + m]]
        return pre, post

    def c_row_layout(self, q):
        nx, nu = self.nx, self.nu
        pat = self.pat_c()[q]
        row = [("x", s) for s in range(nx) if pat[s]] + [("u", c) for c in range(nu) if pat[nx + c]]
        if self.c_t()[q]:
            row += [("T0", 0), ("TF", 0)]
        row += [("a", m) for m in range(self.na) if pat[nx + nu + m]]
        return row

    def tc_row_layout(self, r):
        """Sorted entries of terminal row r as term-var ids (N > 1: x0_s sits at column s*N, xf_s at s*N+N-1)."""
        nx = self.nx
        pat = self.pat_tc()[r]
        order = []
        for s in range(nx):
            order += [nx + s, s]  # x0_s then xf_s
        order += [2 * nx + 1, 2 * nx]  # T0 then TF
        order += [2 * nx + 2 + m for m in range(self.na)]
        return [v for v in order if pat[v]]

    # ---- code generation
    def _var_ref(self):
        ref = {v.name: f"x[{i}]" for i, v in enumerate(self.x)}
        ref.update({v.name: f"u[{i}]" for i, v in enumerate(self.u)})
        ref.update({v.name: f"a[{i}]" for i, v in enumerate(self.a)})
        ref[self.t.name] = "t"
        return ref

    def _term_ref(self):
        ref = {v.name: f"xf[{i}]" for i, v in enumerate(self.xf)}
        ref.update({v.name: f"x0[{i}]" for i, v in enumerate(self.x0)})
        ref.update({v.name: f"a[{i}]" for i, v in enumerate(self.a)})
        ref[self.tf.name], ref[self.t0.name] = "tf", "t0"
        return ref

    @staticmethod
    def _switch(name, values, default=-1):
        cases = " ".join(f"case {i}: return {int(v)};" for i, v in enumerate(values))
        return f"  MPX_HD static constexpr int {name}(int i) {{ switch (i) {{ {cases} default: return {default}; }} }}"

    def cuda_struct(self, name):
        nx, nu, na, nc = self.nx, self.nu, self.na, self.nc
        L = [f"struct {name} {{"]
        L.append(f"  static constexpr int NX = {nx}, NU = {nu}, NA = {na}, NC = {nc}, NTC = {self.ntc};")
        L.append(f"  static constexpr int NJF = {len(self.jf)}, NJC = {len(self.jc)}, NGL = {len(self.gL)}, "
                 f"NJTC = {len(self.jtc)}, NGM = {len(self.gM)};")
        any_t = lambda ds: int(any(not d.is_value(0.0) for d in ds))
        L.append(f"  static constexpr bool F_T = {any_t(self.ft)}, C_T = {any_t(self.ct)}, "
                 f"L_T = {any_t([self.Lt])}, L_NZ = {int(self.L_nz)};")
        # -- F rows
        npre, next_, tpos, diag = [], [], [], []
        pos_of = {}
        for s in range(nx):
            pre, post = self.f_row_layout(s)
            npre.append(len(pre))
            next_.append(len(pre) + len(post))
            tpos.append(len(pre) + post.index(("T0", 0)) if ("T0", 0) in post else -1)
            for i, k in enumerate(pre + post):
                pos_of[(s, k)] = i
        jf_row, jf_pos, jf_var = [], [], []
        for r, v, _ in self.jf:
            jf_row.append(r)
            jf_var.append(v)
            kind = ("x", v) if v < nx else (("u", v - nx) if v < nx + nu else ("a", v - nx - nu))
            jf_pos.append(-1 if kind == ("x", r) else pos_of[(r, kind)])
        for s in range(nx):
            diag.append(next((e for e, (r, v, _) in enumerate(self.jf) if r == s and v == s), -1))
        L.append(self._switch("f_nz", [int(b) for b in self.f_nz], 0))
        L.append(self._switch("f_npre", npre, 0))
        L.append(self._switch("f_next", next_, 0))
        L.append(self._switch("f_tpos", tpos))
        L.append(self._switch("f_diag", diag))
        L.append(self._switch("f_t", self.f_t(), 0))
        # (row s, node variable v) -> entry of jf or -1; column blocks / parameter entries of a mid-point residual row
        # of the adaptive NLP (mpx_adapt_kernel)
        nvv = nx + nu + na
        jf_at = {(r, v): e for e, (r, v, _) in enumerate(self.jf)}
        L.append(self._switch("jf_index", [jf_at.get((s, v), -1) for s in range(nx) for v in range(nvv)]))
        L.append(self._switch("res_nblk", [sum(1 for v in range(nx + nu) if v == s or (s, v) in jf_at)
                                           for s in range(nx)], 0))
        L.append(self._switch("res_na", [sum(1 for m in range(na) if (s, nx + nu + m) in jf_at) for s in range(nx)], 0))
        first = lambda rows, n: [next((e for e, r in enumerate(rows) if r == k), 0) for k in range(n)]
        count = lambda rows, n: [sum(1 for r in rows if r == k) for k in range(n)]
        L.append(self._switch("jf_first", first(jf_row, nx), 0))
        L.append(self._switch("jf_count", count(jf_row, nx), 0))
        L.append(self._switch("jf_row", jf_row))
        L.append(self._switch("jf_pos", jf_pos))
        L.append(self._switch("jf_var", jf_var))
        # -- path rows
        c_len, c_tpos, jc_row, jc_pos, jc_var = [], [], [], [], []
        cpos = {}
        for q in range(nc):
            row = self.c_row_layout(q)
            c_len.append(len(row))
            c_tpos.append(row.index(("T0", 0)) if ("T0", 0) in row else -1)
            for i, k in enumerate(row):
                cpos[(q, k)] = i
        for r, v, _ in self.jc:
            kind = ("x", v) if v < nx else (("u", v - nx) if v < nx + nu else ("a", v - nx - nu))
            jc_row.append(r), jc_var.append(v), jc_pos.append(cpos[(r, kind)])
        L.append(self._switch("c_len", c_len, 0))
        L.append(self._switch("c_tpos", c_tpos))
        L.append(self._switch("jc_first", first(jc_row, nc), 0))
        L.append(self._switch("jc_count", count(jc_row, nc), 0))
        L.append(self._switch("jc_row", jc_row))
        L.append(self._switch("jc_pos", jc_pos))
        L.append(self._switch("jc_var", jc_var))
        L.append(self._switch("gl_var", [v for v, _ in self.gL]))
        # -- terminal rows
        tc_len, jtc_row, jtc_pos, jtc_var = [], [], [], []
        tpos_ = {}
        for r in range(self.ntc):
            row = self.tc_row_layout(r)
            tc_len.append(len(row))
            for i, v in enumerate(row):
                tpos_[(r, v)] = i
        for r, v, _ in self.jtc:
            jtc_row.append(r), jtc_var.append(v), jtc_pos.append(tpos_[(r, v)])
        L.append(self._switch("tc_len", tc_len, 0))
        L.append(self._switch("jtc_row", jtc_row))
        L.append(self._switch("jtc_pos", jtc_pos))
        L.append(self._switch("jtc_var", jtc_var))
        L.append(self._switch("gm_var", [v for v, _ in self.gM]))
        # -- Lagrangian Hessian
        hl = self.hess_layout()
        L.append(f"  static constexpr int NHW = {len(self.hw)}, NHT = {len(self.ht)}, NH_YY = {hl['n_yy']}, "
                 f"NH_AY = {hl['n_ay']}, NH_TY = {len(hl['ty'])}, NH_CORNER = {hl['n_corner']};")
        L.append(self._switch("hw_a", [a for a, _, _ in self.hw]))
        L.append(self._switch("hw_b", [b for _, b, _ in self.hw]))
        L.append(self._switch("hw_cat", hl["cat"]))
        L.append(self._switch("hw_slot", hl["slot"]))
        L.append(self._switch("h_ty_var", hl["ty"]))
        L.append(self._switch("ht_a", [a for a, _, _ in self.ht]))
        L.append(self._switch("ht_b", [b for _, b, _ in self.ht]))

        def fn(sig, outs, targets, ref, prefix):
            lines, refs = tr.emit_c(outs, ref, indent="    ", prefix=prefix)
            body = [f"  __device__ __forceinline__ static void {sig} {{"] + lines
            body += [f"    {t} = {r};" for t, r in zip(targets, refs)]
            body.append("  }")
            return body

        node_sig = "const double* __restrict__ x, const double* __restrict__ u, const double t, const double* __restrict__ a"
        outs = self.f + [d for _, _, d in self.jf] + self.ft
        tg = [f"f[{s}]" for s in range(nx)] + [f"jf[{e}]" for e in range(len(self.jf))] + [f"ft[{s}]" for s in range(nx)]
        L += fn(f"dyn({node_sig}, double* __restrict__ f, double* __restrict__ jf, double* __restrict__ ft)", outs, tg,
                self._var_ref(), "d")
        # one function per row: the row-block kernels evaluate only the row they assemble; jr[] holds the row's
        # packed partials (entries jf_first(s) .. jf_first(s)+jf_count(s)-1 of the full list)
        for s in range(nx):
            ent = [d for (r, _, d) in self.jf if r == s]
            outs = [self.f[s]] + ent + [self.ft[s]]
            tg = ["f[0]"] + [f"jr[{i}]" for i in range(len(ent))] + ["f[1]"]
            L += fn(f"dyn_row(mpx_int<{s}>, {node_sig}, double* __restrict__ f, double* __restrict__ jr)", outs, tg,
                    self._var_ref(), "d")
        outs = self.c + [d for _, _, d in self.jc] + self.ct
        tg = [f"c[{q}]" for q in range(nc)] + [f"jc[{e}]" for e in range(len(self.jc))] + [f"ct[{q}]" for q in range(nc)]
        L += fn(f"path({node_sig}, double* __restrict__ c, double* __restrict__ jc, double* __restrict__ ct)", outs, tg,
                self._var_ref(), "c")
        for q in range(nc):
            ent = [d for (r, _, d) in self.jc if r == q]
            outs = [self.c[q]] + ent + [self.ct[q]]
            tg = ["c[0]"] + [f"jr[{i}]" for i in range(len(ent))] + ["c[1]"]
            L += fn(f"path_row(mpx_int<{q}>, {node_sig}, double* __restrict__ c, double* __restrict__ jr)", outs, tg,
                    self._var_ref(), "c")
        outs = [self.L] + [d for _, d in self.gL] + [self.Lt]
        tg = ["L[0]"] + [f"gl[{e}]" for e in range(len(self.gL))] + ["L[1]"]
        L += fn(f"cost({node_sig}, double* __restrict__ L, double* __restrict__ gl)", outs, tg, self._var_ref(), "q")
        term_sig = ("const double* __restrict__ xf, const double tf, const double* __restrict__ x0, const double t0, "
                    "const double* __restrict__ a")
        outs = self.tc + [d for _, _, d in self.jtc] + [self.M] + [d for _, d in self.gM]
        tg = ([f"tc[{r}]" for r in range(self.ntc)] + [f"jtc[{e}]" for e in range(len(self.jtc))] + ["M[0]"]
              + [f"gm[{e}]" for e in range(len(self.gM))])
        L += fn(f"term({term_sig}, double* __restrict__ tc, double* __restrict__ jtc, double* __restrict__ M, "
                f"double* __restrict__ gm)", outs, tg, self._term_ref(), "m")
        ref = self._var_ref()
        ref[self.hs.name], ref[self.sw.name] = "h", "sw"
        ref.update({v.name: f"lf[{i}]" for i, v in enumerate(self.lamF)})
        ref.update({v.name: f"lc[{i}]" for i, v in enumerate(self.lamC)})
        ref.update({v.name: f"sxs[{i}]" for i, v in enumerate(self.sxs)})
        L += fn(f"hess_node({node_sig}, const double h, const double sw, const double* __restrict__ lf, "
                f"const double* __restrict__ lc, const double* __restrict__ sxs, double* __restrict__ hw)",
                [d for _, _, d in self.hw], [f"hw[{e}]" for e in range(len(self.hw))], ref, "w")
        # value | gradient | Hessian of psi = sum_s lf_s sxs_s f_s  (mid-point residual rows of the adaptive NLP)
        L.append(f"  static constexpr int NRG = {len(self.rg)}, NRH = {len(self.rh)};")
        L.append(self._switch("rg_var", [v for v, _ in self.rg]))
        L.append(self._switch("rh_a", [a for a, _, _ in self.rh]))
        L.append(self._switch("rh_b", [b for _, b, _ in self.rh]))
        L += fn(f"hess_res({node_sig}, const double* __restrict__ lf, const double* __restrict__ sxs, "
                f"double* __restrict__ r)",
                [self.psi] + [d for _, d in self.rg] + [d for _, _, d in self.rh],
                [f"r[{e}]" for e in range(1 + len(self.rg) + len(self.rh))], ref, "r")
        ref = self._term_ref()
        ref[self.sw.name] = "sw"
        ref.update({v.name: f"lt[{i}]" for i, v in enumerate(self.lamT)})
        L += fn(f"hess_term({term_sig}, const double sw, const double* __restrict__ lt, double* __restrict__ ht)",
                [d for _, _, d in self.ht], [f"ht[{e}]" for e in range(len(self.ht))], ref, "z")
        L.append("};")
        return "\n".join(L)


class Program:
    """All phases of one OCP, traced."""

    def __init__(self, ocp):
        self.nx, self.nu, self.na, self.n_phases = ocp.nx, ocp.nu, ocp.na, ocp.n_phases
        self.phases = [PhaseProgram(ocp, ph) for ph in range(ocp.n_phases)]
        self._src = None
        self.cuda_source()
        tr.Expr.reset_interning()  # the traced DAG is finished: do not let the intern table grow with every OCP

    def cuda_source(self) -> str:
        """Canonical generated source: one struct per phase, named by position only (so the hash is stable)."""
        if self._src is None:
            parts = ["// generated by mpopt_b200.program -- node functors traced from the user's Python callables"]
            for k, ph in enumerate(self.phases):
                parts.append(ph.cuda_struct(f"MPX_PHASE_NAME({k})"))
            self._src = "\n".join(parts) + "\n"
        return self._src

    def key(self) -> str:
        return hashlib.sha256(self.cuda_source().encode()).hexdigest()[:16]
