import json
import numpy

import plotContactMaps as plotModule


class Bin:
    def __init__(self, ID, chrom, rowSum):
        self.ID, self.chrom, self.rowSum = ID, chrom, rowSum


def _unbound(*a, **k):
    raise RuntimeError("hot-path function was not rebound by the launcher")


convertMatrix = averageClusterNodes = reorderMatrix = _unbound
pre_process_all_matrix_breakpoints = filter_noisy_breakpoints = _unbound


def runPipeline(matrix_npy, out_json, minSize, modularity, psig):
    A = numpy.load(matrix_npy)
    binList = [Bin(i + 1, "s", float(sum(numpy.asarray(r)))) for i, r in enumerate(A.tolist())]
    adjMat = convertMatrix(A.tolist(), binList, distance=True, similarity=False)
    labels = [b.chrom + '_' + str(b.ID) for b in binList]
    dendrogram = averageClusterNodes(adjMat, labels, noPlot=True)
    adjMat, binList = reorderMatrix(numpy.asmatrix(adjMat), binList, dendrogram['leaves'])
    plotModule.plotContactMap(adjMat)
    adjMat = convertMatrix(adjMat, binList, distance=False, similarity=True)
    argsorted_adjMat = numpy.asarray(numpy.argsort(adjMat, axis=1)[:, ::-1])
    initial = pre_process_all_matrix_breakpoints(argsorted_adjMat, min_size=minSize, min_frac=modularity, psig=psig)
    cuts = filter_noisy_breakpoints(argsorted_adjMat, initial, psig=psig)
    json.dump({"leaves": [int(v) for v in dendrogram['leaves']], "ivl": dendrogram['ivl'][:5],
               "initial": [int(v) for v in initial], "cuts": [int(v) for v in cuts],
               "S_checksum": float(numpy.asarray(adjMat).sum()),
               "R_row0": [int(v) for v in argsorted_adjMat[0][:10]]}, open(out_json, "w"))
