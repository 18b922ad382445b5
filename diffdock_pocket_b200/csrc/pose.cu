// Fused per-step pose update: score -> perturbation, side-chain torsion chain, rigid move, ligand
// torsion chain, Kabsch re-alignment.  One CTA per sample; the chains are sequential by definition
// (each rotation axis depends on the previous rotations), the atoms of a rotation are parallel.
// Arithmetic mirrors the reference's mixed precision: fp32 positions, rotation matrices from
// scipy's Rotation.from_rotvec in float64 (utils/torsion.py:84-88, 268-271), fp32 quaternion path for
// the rigid rotation (utils/geometry.py:39-86), Kabsch rotation via Horn's quaternion eigenproblem
// in float64 (same optimum as the SVD + reflection fix of utils/geometry.py:232-238).
#include "ddp_common.cuh"

namespace {

__device__ void rotvec_to_matrix(const double rv[3], double R[9]) {
    // scipy.spatial.transform.Rotation.from_rotvec(...).as_matrix()
    const double a2 = rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2];
    const double angle = sqrt(a2);
    double scale;
    if (angle <= 1e-3) scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
    else scale = sin(angle / 2.0) / angle;
    double x = scale * rv[0], y = scale * rv[1], z = scale * rv[2], w = cos(angle / 2.0);
    const double n = sqrt(x * x + y * y + z * z + w * w);
    x /= n; y /= n; z /= n; w /= n;
    const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = x2 - y2 - z2 + w2; R[1] = 2 * (xy - zw);      R[2] = 2 * (xz + yw);
    R[3] = 2 * (xy + zw);      R[4] = -x2 + y2 - z2 + w2; R[5] = 2 * (yz - xw);
    R[6] = 2 * (xz - yw);      R[7] = 2 * (yz + xw);      R[8] = -x2 - y2 + z2 + w2;
}

// rotate atoms about the axis pos[u]-pos[v] through pos[v] by theta (fp32 rotvec, fp64 matrix)
__device__ __forceinline__ void bond_rotation(const float *pu, const float *pv, float theta, double R[9]) {
    const float ax = pu[0] - pv[0], ay = pu[1] - pv[1], az = pu[2] - pv[2];
    const float nrm = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(ax, ax), __fmul_rn(ay, ay)), __fmul_rn(az, az)));
    const double rv[3] = {(double)__fdiv_rn(__fmul_rn(ax, theta), nrm), (double)__fdiv_rn(__fmul_rn(ay, theta), nrm),
                          (double)__fdiv_rn(__fmul_rn(az, theta), nrm)};
    rotvec_to_matrix(rv, R);
}

__device__ __forceinline__ void apply_rotation(float *p, const float pv[3], const double R[9]) {
    const double dx = (double)(p[0] - pv[0]), dy = (double)(p[1] - pv[1]), dz = (double)(p[2] - pv[2]);
    p[0] = (float)(dx * R[0] + dy * R[1] + dz * R[2] + (double)pv[0]);
    p[1] = (float)(dx * R[3] + dy * R[4] + dz * R[5] + (double)pv[1]);
    p[2] = (float)(dx * R[6] + dy * R[7] + dz * R[8] + (double)pv[2]);
}

__device__ void axis_angle_to_matrix_f32(float ax, float ay, float az, float R[9]) {
    const float angle = sqrtf(ax * ax + ay * ay + az * az);
    const float half = 0.5f * angle;
    const float s = (fabsf(angle) < 1e-6f) ? (0.5f - (angle * angle) / 48.f) : (sinf(half) / angle);
    const float r = cosf(half), i = ax * s, j = ay * s, k = az * s;
    const float two_s = 2.0f / (r * r + i * i + j * j + k * k);
    R[0] = 1 - two_s * (j * j + k * k); R[1] = two_s * (i * j - k * r);     R[2] = two_s * (i * k + j * r);
    R[3] = two_s * (i * j + k * r);     R[4] = 1 - two_s * (i * i + k * k); R[5] = two_s * (j * k - i * r);
    R[6] = two_s * (i * k - j * r);     R[7] = two_s * (j * k + i * r);     R[8] = 1 - two_s * (i * i + j * j);
}

// largest-eigenvalue eigenvector of a symmetric 4x4 (cyclic Jacobi, float64)
__device__ void max_eigvec4(double A[4][4], double q[4]) {
    double V[4][4] = {{1, 0, 0, 0}, {0, 1, 0, 0}, {0, 0, 1, 0}, {0, 0, 0, 1}};
    for (int sweep = 0; sweep < 30; ++sweep) {
        double off = 0;
        for (int p = 0; p < 4; ++p)
            for (int r = p + 1; r < 4; ++r) off += A[p][r] * A[p][r];
        if (off < 1e-30) break;
        for (int p = 0; p < 3; ++p)
            for (int r = p + 1; r < 4; ++r) {
                if (fabs(A[p][r]) < 1e-300) continue;
                const double theta = (A[r][r] - A[p][p]) / (2.0 * A[p][r]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
                for (int k = 0; k < 4; ++k) {
                    const double akp = A[k][p], akr = A[k][r];
                    A[k][p] = cs * akp - sn * akr; A[k][r] = sn * akp + cs * akr;
                }
                for (int k = 0; k < 4; ++k) {
                    const double apk = A[p][k], ark = A[r][k];
                    A[p][k] = cs * apk - sn * ark; A[r][k] = sn * apk + cs * ark;
                }
                for (int k = 0; k < 4; ++k) {
                    const double vkp = V[k][p], vkr = V[k][r];
                    V[k][p] = cs * vkp - sn * vkr; V[k][r] = sn * vkp + cs * vkr;
                }
            }
    }
    int best = 0;
    for (int k = 1; k < 4; ++k) if (A[k][k] > A[best][best]) best = k;
    for (int k = 0; k < 4; ++k) q[k] = V[k][best];
}

constexpr int kPoseThreads = 64;
constexpr int kMaxLigAtoms = 512;

__global__ void __launch_bounds__(kPoseThreads)
pose_update_kernel(ddp_pose_t P, ddp_step_coef_t C_host, const ddp_step_coef_t *__restrict__ C_dev) {
    const ddp_step_coef_t C = C_dev ? *C_dev : C_host;
    __shared__ float s_flex[kMaxLigAtoms * 3];
    __shared__ float s_rigid[kMaxLigAtoms * 3];
    __shared__ double s_R[9];
    __shared__ float s_pv[3];
    __shared__ float s_misc[16];
    __shared__ double s_acc[kPoseThreads][9];
    const int s = blockIdx.x, tid = threadIdx.x;

    // ---------------- side chains (utils/diffusion_utils.py:63-70, utils/torsion.py:251-278) ----------
    if (P.sc_ptr != nullptr) {
        for (int b = P.sc_ptr[s]; b < P.sc_ptr[s + 1]; ++b) {
            const float theta = __fadd_rn(__fmul_rn(C.a_sc, P.sc_score[b]), __fmul_rn(C.b_sc, P.sc_z ? P.sc_z[b] : 0.f));
            if (theta == 0.f) continue;               // uniform across the block
            __syncthreads();
            if (tid == 0) {
                const int u = P.sc_bonds[2 * b], v = P.sc_bonds[2 * b + 1];
                double R[9];
                bond_rotation(P.atom_pos + 3 * u, P.atom_pos + 3 * v, theta, R);
                for (int k = 0; k < 9; ++k) s_R[k] = R[k];
                s_pv[0] = P.atom_pos[3 * v]; s_pv[1] = P.atom_pos[3 * v + 1]; s_pv[2] = P.atom_pos[3 * v + 2];
            }
            __syncthreads();
            for (int m = P.sc_sub_ptr[b] + tid; m < P.sc_sub_ptr[b + 1]; m += kPoseThreads)
                apply_rotation(P.atom_pos + 3 * P.sc_sub[m], s_pv, s_R);
        }
        __syncthreads();
    }

    // ---------------- ligand (utils/diffusion_utils.py:37-60) -----------------------------------------
    const int a0 = P.lig_ptr[s], na = P.lig_ptr[s + 1] - a0;
    if (na <= 0) return;
    if (na > kMaxLigAtoms) __trap();             // the host checks ddp_pose_max_ligand_atoms(); never skip a ligand silently
    float *pos = P.lig_pos + 3 * (size_t)a0;
    if (tid == 0) {
        float cx = 0.f, cy = 0.f, cz = 0.f;
        for (int i = 0; i < na; ++i) { cx += pos[3 * i]; cy += pos[3 * i + 1]; cz += pos[3 * i + 2]; }
        s_misc[0] = cx / (float)na; s_misc[1] = cy / (float)na; s_misc[2] = cz / (float)na;
        const float rx = __fadd_rn(__fmul_rn(C.a_rot, P.rot_score[3 * s]), __fmul_rn(C.b_rot, P.rot_z ? P.rot_z[3 * s] : 0.f));
        const float ry = __fadd_rn(__fmul_rn(C.a_rot, P.rot_score[3 * s + 1]), __fmul_rn(C.b_rot, P.rot_z ? P.rot_z[3 * s + 1] : 0.f));
        const float rz = __fadd_rn(__fmul_rn(C.a_rot, P.rot_score[3 * s + 2]), __fmul_rn(C.b_rot, P.rot_z ? P.rot_z[3 * s + 2] : 0.f));
        axis_angle_to_matrix_f32(rx, ry, rz, s_misc + 3);
        for (int k = 0; k < 3; ++k)
            s_misc[12 + k] = __fadd_rn(__fmul_rn(C.a_tr, P.tr_score[3 * s + k]), __fmul_rn(C.b_tr, P.tr_z ? P.tr_z[3 * s + k] : 0.f));
    }
    __syncthreads();
    for (int i = tid; i < na; i += kPoseThreads) {
        const float dx = pos[3 * i] - s_misc[0], dy = pos[3 * i + 1] - s_misc[1], dz = pos[3 * i + 2] - s_misc[2];
        const float *R = s_misc + 3;
        for (int k = 0; k < 3; ++k) {
            const float v = dx * R[3 * k] + dy * R[3 * k + 1] + dz * R[3 * k + 2];
            const float r = __fadd_rn(__fadd_rn(v, s_misc[12 + k]), s_misc[k]);
            s_rigid[3 * i + k] = r;
            s_flex[3 * i + k] = r;
        }
    }
    __syncthreads();
    const int t0 = P.tor_ptr ? P.tor_ptr[s] : 0, t1 = P.tor_ptr ? P.tor_ptr[s + 1] : 0;
    if (t1 == t0) {
        for (int i = tid; i < 3 * na; i += kPoseThreads) pos[i] = s_rigid[i];
        return;
    }
    // torsion chain (utils/torsion.py:68-94)
    for (int t = t0; t < t1; ++t) {
        const float theta = __fadd_rn(__fmul_rn(C.a_tor, P.tor_score[t]), __fmul_rn(C.b_tor, P.tor_z ? P.tor_z[t] : 0.f));
        if (theta == 0.f) continue;
        __syncthreads();
        if (tid == 0) {
            const int u = P.tor_bonds[2 * t] - a0, v = P.tor_bonds[2 * t + 1] - a0;
            double R[9];
            bond_rotation(s_flex + 3 * u, s_flex + 3 * v, theta, R);
            for (int k = 0; k < 9; ++k) s_R[k] = R[k];
            s_pv[0] = s_flex[3 * v]; s_pv[1] = s_flex[3 * v + 1]; s_pv[2] = s_flex[3 * v + 2];
        }
        __syncthreads();
        const uint8_t *mask = P.mask_rotate + P.mask_ptr[t];
        for (int i = tid; i < na; i += kPoseThreads)
            if (mask[i]) apply_rotation(s_flex + 3 * i, s_pv, s_R);
    }
    __syncthreads();
    // Kabsch: R, t with R flex + t ~ rigid (utils/geometry.py:209-243)
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    {
        double ca[3] = {0, 0, 0}, cb[3] = {0, 0, 0};
        for (int i = 0; i < na; ++i)
            for (int k = 0; k < 3; ++k) { ca[k] += s_flex[3 * i + k]; cb[k] += s_rigid[3 * i + k]; }
        for (int k = 0; k < 3; ++k) { ca[k] /= na; cb[k] /= na; }
        for (int i = tid; i < na; i += kPoseThreads)
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b)
                    acc[3 * a + b] += ((double)s_flex[3 * i + a] - ca[a]) * ((double)s_rigid[3 * i + b] - cb[b]);
        for (int k = 0; k < 9; ++k) s_acc[tid][k] = acc[k];
        __syncthreads();
        if (tid == 0) {
            double S[9];
            for (int k = 0; k < 9; ++k) { S[k] = 0; for (int w = 0; w < kPoseThreads; ++w) S[k] += s_acc[w][k]; }
            const double Sxx = S[0], Sxy = S[1], Sxz = S[2], Syx = S[3], Syy = S[4], Syz = S[5], Szx = S[6], Szy = S[7], Szz = S[8];
            double N[4][4] = {{Sxx + Syy + Szz, Syz - Szy, Szx - Sxz, Sxy - Syx},
                              {Syz - Szy, Sxx - Syy - Szz, Sxy + Syx, Szx + Sxz},
                              {Szx - Sxz, Sxy + Syx, -Sxx + Syy - Szz, Syz + Szy},
                              {Sxy - Syx, Szx + Sxz, Syz + Szy, -Sxx - Syy + Szz}};
            double q[4];
            max_eigvec4(N, q);
            const double w = q[0], x = q[1], y = q[2], z = q[3];
            double R[9] = {w * w + x * x - y * y - z * z, 2 * (x * y - w * z), 2 * (x * z + w * y),
                           2 * (x * y + w * z), w * w - x * x + y * y - z * z, 2 * (y * z - w * x),
                           2 * (x * z - w * y), 2 * (y * z + w * x), w * w - x * x - y * y + z * z};
            for (int k = 0; k < 9; ++k) s_R[k] = R[k];
            for (int k = 0; k < 3; ++k)
                s_misc[k] = (float)(-(R[3 * k] * ca[0] + R[3 * k + 1] * ca[1] + R[3 * k + 2] * ca[2]) + cb[k]);
        }
        __syncthreads();
    }
    for (int i = tid; i < na; i += kPoseThreads) {
        const double fx = s_flex[3 * i], fy = s_flex[3 * i + 1], fz = s_flex[3 * i + 2];
        for (int k = 0; k < 3; ++k)
            pos[3 * i + k] = (float)(fx * s_R[3 * k] + fy * s_R[3 * k + 1] + fz * s_R[3 * k + 2]) + s_misc[k];
    }
}

}  // namespace

extern "C" int ddp_pose_max_ligand_atoms(void) { return kMaxLigAtoms; }

extern "C" int ddp_pose_update(const ddp_pose_t *pose, const ddp_step_coef_t *coef, void *stream) {
    if (!pose || !coef) return DDP_E_ARG;
    const ddp_pose_t &P = *pose;
    if (!P.lig_pos || !P.lig_ptr || !P.tr_score || !P.rot_score) return DDP_E_ARG;
    if (P.tor_ptr && (!P.tor_bonds || !P.mask_rotate || !P.mask_ptr || !P.tor_score)) return DDP_E_ARG;
    if (P.sc_ptr && (!P.sc_bonds || !P.sc_sub_ptr || !P.sc_sub || !P.sc_score || !P.atom_pos)) return DDP_E_ARG;
    if (P.n_samples <= 0) return 0;
    pose_update_kernel<<<P.n_samples, kPoseThreads, 0, (cudaStream_t)stream>>>(P, *coef, nullptr);
    DDP_LAUNCH_CHECK();
    return 0;
}

extern "C" int ddp_pose_update_dev(const ddp_pose_t *pose, const ddp_step_coef_t *coef_dev, void *stream) {
    if (!pose || !coef_dev) return DDP_E_ARG;
    const ddp_pose_t &P = *pose;
    if (!P.lig_pos || !P.lig_ptr || !P.tr_score || !P.rot_score) return DDP_E_ARG;
    if (P.tor_ptr && (!P.tor_bonds || !P.mask_rotate || !P.mask_ptr || !P.tor_score)) return DDP_E_ARG;
    if (P.sc_ptr && (!P.sc_bonds || !P.sc_sub_ptr || !P.sc_sub || !P.sc_score || !P.atom_pos)) return DDP_E_ARG;
    if (P.n_samples <= 0) return 0;
    ddp_step_coef_t dummy = {0, 0, 0, 0, 0, 0, 0, 0};
    pose_update_kernel<<<P.n_samples, kPoseThreads, 0, (cudaStream_t)stream>>>(P, dummy, coef_dev);
    DDP_LAUNCH_CHECK();
    return 0;
}
