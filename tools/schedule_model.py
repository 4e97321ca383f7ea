"""Happens-before model of the critical-first ("split") schedule of vc_chol_factor (csrc/vc_chol.cu, panel_split and
the lookahead loop), host calls mirrored one for one.
Each kernel: stream, reads (set of tiles), writes (set), reds (set of tiles updated by RED).
Edges: stream order + event record/wait.  Check: every conflicting pair is ordered."""
import itertools, sys

class Sim:
    def __init__(self):
        self.k = []            # kernels: dict(name, stream, R, W, A)
        self.last = {}         # stream -> last node id
        self.pred = {}         # node -> set(preds)
        self.ev = {}           # event -> node id at record time (or None)
        self.pending_waits = {}  # stream -> set(nodes) to be added as preds of the next node in that stream
    def _node(self, stream, name, R=(), W=(), A=()):
        i = len(self.k)
        self.k.append(dict(name=name, stream=stream, R=set(R), W=set(W), A=set(A)))
        p = set()
        if stream in self.last: p.add(self.last[stream])
        p |= self.pending_waits.pop(stream, set())
        self.pred[i] = p
        self.last[stream] = i
        return i
    def launch(self, stream, name, R=(), W=(), A=()): return self._node(stream, name, R, W, A)
    def record(self, ev, stream):
        # an event captures everything before it in the stream: model as a marker node
        self.ev[ev] = self._node(stream, "rec:" + ev)
    def wait(self, stream, ev):
        if self.ev.get(ev) is not None:
            self.pending_waits.setdefault(stream, set()).add(self.ev[ev])
    def closure(self):
        n = len(self.k)
        anc = [set() for _ in range(n)]
        for i in range(n):
            for p in self.pred[i]:
                anc[i].add(p); anc[i] |= anc[p]
        return anc
    def check(self):
        anc = self.closure()
        bad = []
        real = [i for i, k in enumerate(self.k) if not k["name"].startswith("rec:")]
        for a, b in itertools.combinations(real, 2):
            if a in anc[b] or b in anc[a]: continue
            ka, kb = self.k[a], self.k[b]
            conf = (ka["W"] & (kb["R"] | kb["W"] | kb["A"])) | (kb["W"] & (ka["R"] | ka["W"] | ka["A"])) \
                 | (ka["A"] & kb["R"]) | (kb["A"] & ka["R"])
            if conf: bad.append((ka["name"], kb["name"], sorted(conf)[:3]))
        return bad

def trapezoid(J0, J1, nrows):
    return [(I, J) for J in range(J0, J1) for I in range(J, nrows)]   # order irrelevant except first = (J0, J0)

def factor_split(nt, nb, nbt=1, nA_frac=None, evals=1, drop=None):
    S = Sim()
    nrows = nt + nbt
    sm, sp, sq = "sm", "sp", "sq"
    LIN = lambda j: ("linv", j)
    for e in range(evals):
        S.launch(sm, "build%d" % e, W=[(I, J) for J in range(nt) for I in range(J, nrows)] + [LIN(j) for j in range(0)])
        S.launch(sm, "reset")
        S.record("ev_update", sm); S.wait(sp, "ev_update")
        def panel_split(ob, wait_ev):
            oe = min(ob + nb, nt)
            pending = wait_ev
            if wait_ev and drop != "sq_ev_col": S.wait(sq, wait_ev)
            for jt in range(ob, oe):
                S.launch(sp, "potrf%d" % jt, R=[(jt, jt)], W=[(jt, jt), LIN(jt)])
                S.record("ev_p", sp); S.wait(sq, "ev_p")
                if pending:
                    if not (drop == "sp_pending"): S.wait(sp, pending)
                    pending = None
                n_main = nt - (jt + 1)
                if n_main > 0:
                    S.launch(sp, "Ta%d" % jt, R=[(jt + 1, jt), LIN(jt)], W=[(jt + 1, jt)])
                rows = list(range(jt + 2, nt)) + list(range(nt, nrows))
                if rows:
                    S.launch(sq, "Tb%d" % jt, R=[(I, jt) for I in rows] + [LIN(jt)], W=[(I, jt) for I in rows])
                tiles = trapezoid(jt + 1, oe, nrows)
                if tiles:
                    f = tiles[0]
                    S.launch(sp, "Ua%d" % jt, R=[(f[0], jt)], A=[f])
                    S.record("ev_t", sp)
                    if drop != "ev_t": S.wait(sq, "ev_t")
                    rest = tiles[1:]
                    if rest:
                        S.launch(sq, "Ub%d" % jt, R=[(I, jt) for I, J in rest] + [(J, jt) for I, J in rest], A=rest)
                    S.record("ev_u", sq); pending = "ev_u"
            S.record("ev_q", sq)
            if drop != "join": S.wait(sp, "ev_q")
        panel_split(0, None)
        S.record("ev_panel", sp)
        for ob in range(0, nt, nb):
            oe = min(ob + nb, nt)
            if oe >= nt: break
            ne = min(oe + nb, nt)
            S.wait(sm, "ev_panel")
            ops = lambda tl: [(I, c) for I, J in tl for c in range(ob, oe)] + [(J, c) for I, J in tl for c in range(ob, oe)]
            col = trapezoid(oe, ne, nrows)
            S.launch(sm, "U0_%d" % ob, R=ops(col[:1]), A=col[:1])
            S.record("ev_c0", sm)
            if drop != "ev_c0": S.wait(sp, "ev_c0")
            if col[1:]: S.launch(sm, "Uc_%d" % ob, R=ops(col[1:]), A=col[1:])
            S.record("ev_col", sm)
            rest = trapezoid(ne, nt, nrows)
            nA = len(rest) if nA_frac is None else int(len(rest) * nA_frac)
            if rest[:nA]: S.launch(sm, "RA_%d" % ob, R=ops(rest[:nA]), A=rest[:nA])
            panel_split(oe, "ev_col")
            S.record("ev_panel", sp)
            if nA < len(rest):
                S.wait(sm, "ev_panel")
                S.launch(sm, "RB_%d" % ob, R=ops(rest[nA:]), A=rest[nA:])
        S.wait(sm, "ev_panel")
        S.launch(sm, "finalize%d" % e, R=[(I, J) for J in range(nt) for I in range(nt, nrows)] + [LIN(j) for j in range(nt)])
    return S

def main():
    for nt, nb in [(5, 1), (6, 2), (7, 3), (9, 2)]:
        for frac in (None, 0.5):
            for evals in (1, 2):
                bad = factor_split(nt, nb, 1, frac, evals).check()
                print("nt=%d nb=%d nA=%s evals=%d: %d unordered conflicts" % (nt, nb, frac, evals, len(bad)))
                for b in bad[:8]: print("   ", b)
    for drop in ("sq_ev_col", "sp_pending", "join", "ev_c0", "ev_t"):
        bad = factor_split(7, 3, 1, 0.5, 1, drop).check()
        print("negative control, dropped wait %-10s: %d unordered conflicts, e.g. %s" % (drop, len(bad), bad[:1]))


if __name__ == "__main__":
    main()
