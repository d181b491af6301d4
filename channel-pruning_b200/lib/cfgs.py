"""Mirror of the part of the reference's global configuration that the pruning hot
path reads (reference lib/cfgs.py).  Same names, same defaults, same mutability:

  cfgs.alpha                 (cfgs.py:18)   start bracket of the alpha search, *carried
                                            across layers* (decompose.py:491, 627)
  cfgs.c / dcfgs.dic.rank_tol (cfgs.py:84)  acceptance window, overrides the argument
                                            (decompose.py:393)
  dcfgs.nBatches / nPointsPerLayer (cfgs.py:104,108)
  dcfgs.autodet, solver, ls, fc_ridge, nonlinear_fc, nofc, dic.alter, dic.debug,
  dic.vh, dic.keep, model, res.short -- only their c3 defaults are implemented; any other
  value raises NotImplementedError at the point where the reference would branch.
"""


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


alpha = 1e-3  # cfgs.py:18


class solvers:  # cfgs.py:41-47
    lightning = 'lightning'
    sk = 'sklearn'
    lowparams = 'lowparams'
    gd = 'gd'
    keras = 'keras'
    tls = "tls"


class Models:  # cfgs.py:58-62
    vgg = 'vgg'
    xception = 'xception'
    resnet = 'resnet'
    rescifar = 'rescifar'


c = _AttrDict()
c.dic = _AttrDict()
c.dic.option = 0
c.dic.layeralpha = 1
c.dic.debug = 0
c.dic.keep = 3.
c.dic.rank_tol = .1
c.dic.alter = 0
c.dic.vh = 1
c.res = _AttrDict()
c.res.short = 0
c.res.bn = 1
c.fc_ridge = 0
c.ls = 'linear'
c.nonlinear_fc = 0
c.nofc = 0
c.nBatches = 500
c.nPointsPerLayer = 10
c.autodet = False
c.solver = solvers.sk
c.model = ''


def set_nBatches(n):  # cfgs.py:119-121
    c.nBatches = n
    c.nBatches_fc = c.nBatches
