"""Render gpurun_out/parity_table.jsonl (written by tests/parity.py during `pytest -m gpu`) as the markdown table of
DESIGN.md §5:  python tools/parity_table.py [gpurun_out/parity_table.jsonl] > profiles/r02_parity_table.md"""
import json
import sys


def main(path):
    seen = {}
    for line in open(path):
        d = json.loads(line)
        seen[d["config"]] = d           # the last run of a configuration wins
    print("| configuration (all against the fp32 CPU oracle / the unmodified reference's golden fixtures) | activations max-rel-err "
          "(worst of D(x), G(z), D(G(z))) | cos D-real | cos D-fake | cos D real+fake | cos G-step | losses (worst rel.) |")
    print("|---|---|---|---|---|---|---|")
    for name, d in seen.items():
        rows = d["rows"]
        act = max((v for k, n, v in rows if k == "act"), default=float("nan"))
        loss = max((v for k, n, v in rows if k == "loss"), default=float("nan"))
        cos = {n: v for k, n, v in rows if k == "cos"}

        def c(n):
            return "%.6f" % cos[n] if n in cos else "—"
        print("| %s | %.1e | %s | %s | %s | %s | %.1e |" % (name, act, c("D-real"), c("D-fake"), c("D-accum"), c("G-step"), loss))
    print()
    print("north_star bars: activations <= 1e-2, every cosine >= 0.999, losses <= 2e-2 (tests/parity.py asserts them on every row).")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_table.jsonl")
