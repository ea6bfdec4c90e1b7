"""B200-native fused alignment loss heads (NT-Xent + DPO)."""
