"""B200-native batched simulator for the DeepMimic-style humanoid env of Galleta12/Physics-Based-RL."""
