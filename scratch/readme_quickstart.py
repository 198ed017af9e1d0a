import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from functools import partial
from golemflavor_b200 import fr, llh, mcmc
from golemflavor_b200.enums import ParamTag, PriorsCateg
from golemflavor_b200.param import Param, ParamSet

measured = fr.u_to_fr(fr.normalize_fr((1, 0, 0)), fr.NUFIT_U)          # injected composition
a1, a2 = fr.fr_to_angles(measured)
asimov = ParamSet([Param('measured_angle1', a1, [0., 1.], std=0.02, tag=ParamTag.BESTFIT),
                   Param('measured_angle2', a2, [-1., 1.], std=0.02, tag=ParamTag.BESTFIT)])
lg, t = PriorsCateg.LIMITEDGAUSS, ParamTag.SM_ANGLES
pset = ParamSet([Param('s_12_2', 0.307, [0., 1.], seed=[0.26, 0.35], std=0.013, prior=lg, tag=t),
                 Param('c_13_4', (1 - 0.02206) ** 2, [0., 1.], seed=[0.950, 0.961], std=0.00147, prior=lg, tag=t),
                 Param('s_23_2', 0.538, [0., 1.], seed=[0.31, 0.75], std=0.069, prior=lg, tag=t),
                 Param('dcp', 4.08404, [0., 2 * np.pi], std=2.0, tag=t),
                 Param('source_angle1', 0, [0., 1.], tag=ParamTag.SRCANGLES),
                 Param('source_angle2', 0, [-1., 1.], tag=ParamTag.SRCANGLES)])
from argparse import Namespace
ln_prob = partial(llh.ln_prob, args=Namespace(source_ratio=[1, 2, 0], no_bsm=True),
                  asimov_paramset=asimov, llh_paramset=pset)
ln_prob([0.31, 0.956, 0.5, 1.0, 0.95, 0.9])            # one point  -> float   (as the reference)
ln_prob(mcmc.flat_seed(pset, 4096))                    # a batch    -> [4096]  (one kernel launch)
samples = mcmc.mcmc(mcmc.flat_seed(pset, 1024), ln_prob, 6, 1024, burnin=100, nsteps=200)  # on-device sampler

print('quick start ok', samples.shape, float(ln_prob([0.31, 0.956, 0.5, 1.0, 0.95, 0.9])))
