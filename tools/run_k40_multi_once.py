import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gp_dla_detection_b200 import api, synthetic as syn
prior = syn.make_prior()
m40 = syn.make_model(40); s = syn.make_samples(10000)
api.DLAProcessor(m40, s, prior).process(syn.make_spectra(m40, 37, seed=1), return_sample_log_likelihoods=False)
m20 = syn.make_model(20); sl = syn.make_samples(10000, with_lls=True)
api.DLAProcessor(m20, sl, prior).process_multi(syn.make_spectra(m20, 37, seed=2, meanflux=True, dla_fraction=0.3), return_samples=False)
