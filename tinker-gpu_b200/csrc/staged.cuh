// Constants of the staged real-space operator (staged.cu).
#pragma once
#define SG_GROUP 64            // atoms per CTA: two 32-atom Morton blocks
#define SG_LANES 4             // lanes per atom
#define SG_THREADS (SG_GROUP * SG_LANES)
#define SG_VJB_CAP 256         // Verlet j-blocks a group may reach (overflow: the row kernels take over)
#define SG_SLOT_MASK 0x7fff
#define SG_LISTED 0x8000u      // the row entry is a listed (excluded / scaled) pair: image of ROW_LISTED_FLAG
