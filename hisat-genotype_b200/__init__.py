"""B200-native typing hot path of HISAT-genotype (stage a: per-read allele compatibility,
stage b: EM abundance).  See DESIGN.md."""
__version__ = "0.1.0"
