# A/B sweeps behind DESIGN.md 4.1 (tuning build: GM3D_NVCC_FLAGS=-DGM3D_TUNING_ENV python -m gm3d_b200.build --force)
for regs in 56 64 72 80; do GM3D_CS_REGS=$regs python tools/quick_step.py --path single; GM3D_CS_REGS=$regs python tools/quick_step.py --path dataflow; done
for nl in 2 3 4 6 8 12 16 24; do GM3D_RING_LANES=$nl python tools/quick_step.py --path dataflow; done
for cw in 4 8; do for t in 1 2 3 4; do GM3D_CD_WARPS=$cw GM3D_CD_TRIPS=$t python tools/quick_step.py --path dataflow; done; done
for parts in g gm c gc gmc; do python tools/quick_step.py --path dataflow --parts $parts; done
