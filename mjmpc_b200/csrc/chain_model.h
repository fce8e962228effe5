// Flat parameter block of a 7-hinge serial chain (doubles).  Produced on the host by
// mjmpc_b200/envs/model.py::_merge_chain (the Python constants CH_* mirror these) from the
// reference's MJCF (mjmpc/envs/assets/xml/sawyer.xml) and consumed by the rollout kernel.
#pragma once

#define MJB_NJ 7

enum {
    CH_OFF = 0,                    // 7*3  link frame origin in parent link frame
    CH_MASS = CH_OFF + 21,         // 7
    CH_COM = CH_MASS + 7,          // 7*3  COM in link frame
    CH_INERTIA = CH_COM + 21,      // 7*6  (xx,yy,zz,xy,xz,yz) about COM, link axes
    CH_ARMATURE = CH_INERTIA + 42, // 7
    CH_DAMPING = CH_ARMATURE + 7,  // 7
    CH_GEAR = CH_DAMPING + 7,      // 7
    CH_CTRL_LO = CH_GEAR + 7,      // 7
    CH_CTRL_HI = CH_CTRL_LO + 7,   // 7
    CH_RANGE_LO = CH_CTRL_HI + 7,  // 7
    CH_RANGE_HI = CH_RANGE_LO + 7, // 7
    CH_INVW0 = CH_RANGE_HI + 7,    // 7    dof_invweight0
    CH_SCALARS = CH_INVW0 + 7,     // 20 scalars, see below
    CH_NPARAM = CH_SCALARS + 20,   // host block ends here (167)
    // derived on upload (mjb_model_create): first moment h = m*com and inertia about the link origin
    CH_H = CH_NPARAM,              // 7*3
    CH_IO = CH_H + 21,             // 7*6 (xx,yy,zz,xy,xz,yz)
    CH_HDAMP = CH_IO + 42,         // 7    timestep * damping
    CH_NDEV = CH_HDAMP + 7         // 237
};

enum {
    CS_TIMESTEP = CH_SCALARS + 0,
    CS_SOLK = CH_SCALARS + 1,      // constraint reference stiffness
    CS_SOLB = CH_SCALARS + 2,      // constraint reference damping
    CS_IMP_D0 = CH_SCALARS + 3,
    CS_IMP_DW = CH_SCALARS + 4,
    CS_IMP_WIDTH = CH_SCALARS + 5,
    CS_IMP_MID = CH_SCALARS + 6,
    CS_IMP_POWER = CH_SCALARS + 7,
    CS_HAND = CH_SCALARS + 8,      // 3
    CS_CON_POS = CH_SCALARS + 11,  // 3
    CS_CON_RADIUS = CH_SCALARS + 14,
    CS_CON_PLANE_Z = CH_SCALARS + 15,
    CS_CON_MARGIN = CH_SCALARS + 16,
    CS_CON_INVW = CH_SCALARS + 17,
    CS_LIMITED_MASK = CH_SCALARS + 18,
    CS_FRAME_SKIP = CH_SCALARS + 19
};

// Fill the derived tail (CH_H, CH_IO, CH_HDAMP) of a CH_NDEV block whose first CH_NPARAM
// entries hold the host block.
static inline void mjb_derive_params(double* P) {
    for (int l = 0; l < MJB_NJ; l++) {
        const double m = P[CH_MASS + l];
        const double cx = P[CH_COM + 3 * l], cy = P[CH_COM + 3 * l + 1], cz = P[CH_COM + 3 * l + 2];
        const double* I = P + CH_INERTIA + 6 * l;
        P[CH_H + 3 * l] = m * cx; P[CH_H + 3 * l + 1] = m * cy; P[CH_H + 3 * l + 2] = m * cz;
        const double cc = cx * cx + cy * cy + cz * cz;
        double* O = P + CH_IO + 6 * l;
        O[0] = I[0] + m * (cc - cx * cx); O[1] = I[1] + m * (cc - cy * cy); O[2] = I[2] + m * (cc - cz * cz);
        O[3] = I[3] - m * cx * cy; O[4] = I[4] - m * cx * cz; O[5] = I[5] - m * cy * cz;
        P[CH_HDAMP + l] = P[CS_TIMESTEP] * P[CH_DAMPING + l];
    }
}

// 1 when the block has exactly the structural zeros SawyerTraits (chain_dynamics.cuh) assumes.
static inline int mjb_params_fit_sawyer(const double* P) {
    static const int offm[7] = {7, 1, 0, 1, 0, 1, 0}, comm[7] = {7, 0, 1, 0, 1, 0, 1};
    for (int l = 0; l < MJB_NJ; l++) {
        for (int k = 0; k < 3; k++) {
            if (!((offm[l] >> k) & 1) && P[CH_OFF + 3 * l + k] != 0.0) return 0;
            if (!((comm[l] >> k) & 1) && P[CH_COM + 3 * l + k] != 0.0) return 0;
        }
        if (l > 0) for (int k = 3; k < 6; k++) if (P[CH_INERTIA + 6 * l + k] != 0.0) return 0;
    }
    return 1;
}
