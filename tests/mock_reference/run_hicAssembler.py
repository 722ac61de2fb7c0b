import argparse
import json

if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("-part1", action="store_true")
    ap.add_argument("-part2", action="store_true")
    ap.add_argument("-config")
    a = ap.parse_args()
    cfg = json.load(open(a.config))
    if a.part1:
        import scaffoldToChromosomes as part1
        part1.runPipeline(cfg["matrix"], cfg["out1"], cfg["minSize"], cfg["modularity"], cfg["psig"])
    if a.part2:
        import orderGenome as part2
        part2.runPipeline(cfg["matrix"], cfg["nodes"], cfg["out2"], cfg["nScaffolds"], cfg["scanScaffolds"])
