"""B200-native (sm_100a) implementation of DiffDock-Pocket's reverse-diffusion score-model hot path.

Public surface (mirrors the reference's interfaces for this path):
  get_model(args, device, t_to_sigma, no_parallel, confidence_mode)   utils/utils.py:59
  TensorProductScoreModel.forward(data)                              models/all_atom_score_model.py:238
  TensorProductConvLayer.forward(...)                                models/score_model.py:108
  sampling(...), randomize_position(...)                             utils/sampling.py:16,70
  radius / radius_graph / knn_graph / scatter                        torch_cluster / torch_scatter
"""
__version__ = '0.1.0'
