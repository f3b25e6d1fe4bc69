"""bayesflux_b200: B200-native engine for the BayesFlux hot path (minibatch log-posterior
gradient of a BNN + the SG-MCMC update), behind a mirror of BayesFlux's API.

The directory is called ``bayesflux.jl_b200``; import it as ``bayesflux_b200`` (see
bayesflux_b200.py at the repository root).  Importing fails loudly when the CUDA engine
(libbfx.so) has not been built: there is no CPU fallback."""
from . import _lib  # noqa: F401  (raises ImportError when libbfx.so is missing)
from ._lib import BfxError, LIB_PATH  # noqa: F401
from .api import *  # noqa: F401,F403
from .api import (BNN, Chain, Dense, LSTM, RNN, NetConstructor, destruct, Gamma, InverseGamma, Normal, Uniform,  # noqa: F401
                  FeedforwardNormal, FeedforwardTDist, SeqToOneNormal, SeqToOneTDist, GaussianPrior,
                  MixtureScalePrior, InitialiseAllSame, split_params, loglikeprior, grad_loglikeprior,
                  ConstantStepsize, DualAveragingStepSize, FixedMassAdapter, DiagCovMassAdapter,
                  RMSPropMassAdapter, FullCovMassAdapter, MCMCState, SGLD, SGNHT, SGNHTS, GGMC, HMC, mcmc,
                  advance, launch_count, comm_unique_id, comm_init, comm_destroy, graph_replays, peer_ranks, Descent, Momentum, RMSProp, ADAM, FluxModeFinder, find_mode, predict,
                  sample_posterior_predict, get_posterior_networks, chain_diagnostics,
                  identity, relu, tanh, sigmoid)
from .sharded import (shard_chains, shard_data, global_num_batches, model_regime, ShardedChain, peer_slice,
                      peer_exchange_reference,  # noqa: F401,E402
                      mcmc_sharded)
