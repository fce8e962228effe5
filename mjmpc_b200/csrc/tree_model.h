// Parameter layout of the runtime-parameterised tree rollout kernel (rollout_tree.cu).  Mirrors
// mjmpc_b200/envs/mjcf_tree.py (pack_links); tests/test_tree_cpu.py checks the two stay in step through
// mjb_tree_layout().
//
// The model is lowered to ONE LINK PER DOF: a body with several joints becomes a run of links, all but the last
// massless.  A link's frame sits at its joint anchor and turns / slides with the joint.
#pragma once

#define MJB_TREE_MAX_LINKS 12
#define MJB_TREE_HINGE 0
#define MJB_TREE_SLIDE 1

// per link, doubles
enum {
    LK_RFIX = 0,     // 9: link frame at q = 0 -> parent link frame
    LK_OFF = 9,      // 3: link origin in the parent link frame (q = 0)
    LK_AXIS = 12,    // 3: joint axis, link frame
    LK_MASS = 15,    // body carried by this link (0 for the massless links of a multi-joint body)
    LK_COM = 16,     // 3
    LK_IC = 19,      // 6: inertia about the centre of mass, link frame: xx yy zz xy xz yz
    LK_RIN = 25,     // 9: rows = axes of the body's inertial frame in link coordinates (fluid model)
    LK_BOX = 34,     // 3: equivalent inertia box (fluid model)
    LK_ARM = 37, LK_DAMP = 38, LK_STIFF = 39, LK_SREF = 40, LK_LO = 41, LK_HI = 42,
    LK_INVW = 43,    // dof_invweight0
    LK_SOLK = 44, LK_SOLB = 45,
    LK_SOLIMP = 46,  // 5
    LK_GEAR = 51, LK_CLO = 52, LK_CHI = 53,
    LK_STRIDE = 54
};
// per link, ints
// LI_BODY: bit 0 = the link carries a body, bit 1 = LK_RFIX is the identity
enum { LI_PARENT = 0, LI_TYPE = 1, LI_LIMITED = 2, LI_ACT = 3, LI_BODY = 4, LI_STRIDE = 5 };
// globals, doubles
enum { TG_DT = 0, TG_GRAV = 1, TG_RHO = 4, TG_VISC = 5, TG_STRIDE = 6 };
