import json
import numpy


class Bin:
    def __init__(self, ID):
        self.ID = ID


class Scaffold:
    def __init__(self, name, binList, orientation):
        self.name, self.binList, self.orientation = name, binList, orientation

    def flipOrientation(self):
        self.orientation = "-" if self.orientation == "+" else "+"
        self.binList = self.binList[::-1]


def _unbound(*a, **k):
    raise RuntimeError("hot-path function was not rebound by the launcher")


costFunction_numba = bruteForceBestScore = checkAllScores = scanOrdering = _unbound


def giveNewAdjMat(matrix, scaffList, binList):
    nList = [n for s in scaffList for n in s.binList]
    orderDict = {g: i for i, g in enumerate(nList)}
    binIndices = {b.ID: i for i, b in enumerate(binList)}
    idx = [binIndices[n] for n in nList]
    return numpy.asmatrix(matrix[numpy.ix_(idx, idx)]), orderDict


def reorderScaffList(orderList, orientationList, scaffDict):
    out = []
    for name, o in zip(orderList, orientationList):
        if scaffDict[name].orientation != o:
            scaffDict[name].flipOrientation()
        out.append(scaffDict[name])
    return out


def runPipeline(matrix_npy, nodes_json, out_json, nScaffolds, scanScaffolds):
    matrix = numpy.asmatrix(numpy.load(matrix_npy))
    nodes = json.load(open(nodes_json))
    binList = [Bin(i + 1) for i in range(matrix.shape[0])]
    scaffDict = {}
    for b, nm in nodes:
        scaffDict.setdefault(nm, Scaffold(nm, [], "+")).binList.append(b)
    pool = sorted(scaffDict.values(), key=lambda s: len(s.binList), reverse=True)
    ordered, pool = pool[:nScaffolds], pool[nScaffolds:]
    adjMat, orderDict = giveNewAdjMat(matrix, ordered, binList)
    order, orient, score = bruteForceBestScore(ordered, scaffDict, adjMat, orderDict)
    ordered = reorderScaffList(order, orient, scaffDict)
    while True:
        if pool:
            ordered.append(pool.pop(0))
        adjMat, orderDict = giveNewAdjMat(matrix, ordered, binList)
        new = ordered.pop(-1)
        ordered, bestCost = checkAllScores(adjMat, orderDict, ordered, new)
        if not pool:
            break
    if len(ordered) > nScaffolds:
        ordered, bestCost = scanOrdering(ordered, scaffDict, orderDict, matrix, binList, bestCost,
                                         scanScaffolds=scanScaffolds)
    adjMat, orderDict = giveNewAdjMat(matrix, ordered, binList)
    total = sum([numpy.trace(adjMat, offset=i) for i in range(1, len(adjMat))])
    json.dump({"ordering": [[s.name, s.orientation] for s in ordered], "bestCost": float(bestCost),
               "finalCost": float(costFunction_numba(adjMat, total))}, open(out_json, "w"))
