// nbx_bonded.cu -- SPC/Fw intramolecular terms, one thread per water molecule (sm_100a).
//
//   harmonic_bond_potential_acceleration!    src/basic_potentials.jl:367-393
//     partner table src/nbody_to_ode.jl:263-288: O -> (H1, H2), H1 -> (O), H2 -> (O); no minimum image
//   valence_angle_potential_acceleration!    src/basic_potentials.jl:395-433   (a = H1, b = O, c = H2)
// Columns of molecule m: O = 3m, H1 = 3m+1, H2 = 3m+2 (src/nbody_to_ode.jl:46).  O(N), HBM-bound.
#include "nbx_internal.cuh"

namespace nbx {

struct V3 { double x, y, z; };
__device__ __forceinline__ V3 sub(const V3 &a, const V3 &b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ double dot(const V3 &a, const V3 &b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ double norm(const V3 &a) { return sqrt(dot(a, a)); }
__device__ __forceinline__ V3 cross(const V3 &a, const V3 &b)
{
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ V3 scale(const V3 &a, double s) { return V3{a.x * s, a.y * s, a.z * s}; }

// -(r - rOH) k / r * rij   (:381-387)
__device__ __forceinline__ V3 bond_force(const V3 &ri, const V3 &rj, double rOH, double kb)
{
    const V3 rij = sub(ri, rj);
    const double r = norm(rij);
    const double factor = -(r - rOH) * kb / r;
    return scale(rij, factor);
}

__global__ void spcfw_kernel(const double *__restrict__ pos, int64_t ld, const double *__restrict__ mass, int mlo,
                             int mhi, double rOH, double aHOH0, double kb, double ka, double *__restrict__ acc)
{
    const int m = mlo + blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= mhi) return;
    const int64_t o = 3 * (int64_t)m, h1 = o + 1, h2 = o + 2;
    const V3 rO{pos[o], pos[ld + o], pos[2 * ld + o]};
    const V3 rA{pos[h1], pos[ld + h1], pos[2 * ld + h1]};
    const V3 rC{pos[h2], pos[ld + h2], pos[2 * ld + h2]};
    const double imO = 1.0 / mass[o], imA = 1.0 / mass[h1], imC = 1.0 / mass[h2];

    // bonds: O sees both hydrogens, each hydrogen sees the oxygen
    const V3 fOA = bond_force(rO, rA, rOH, kb), fOC = bond_force(rO, rC, rOH, kb);
    const V3 fA = bond_force(rA, rO, rOH, kb), fC = bond_force(rC, rO, rOH, kb);
    V3 aO{imO * (fOA.x + fOC.x), imO * (fOA.y + fOC.y), imO * (fOA.z + fOC.z)};
    V3 aA = scale(fA, imA), aC = scale(fC, imC);

    // angle (:404-432)
    const V3 rba = sub(rA, rO), rbc = sub(rC, rO), rcb = sub(rO, rC);
    const V3 X = cross(rba, rbc);
    V3 pa = cross(rba, X), pc = cross(rcb, X);
    pa = scale(pa, 1.0 / norm(pa));
    pc = scale(pc, 1.0 / norm(pc));
    const double nba = norm(rba), nbc = norm(rbc);
    double cosine = dot(rba, rbc) / (nba * nbc);
    cosine = cosine > 1.0 ? 1.0 : (cosine < -1.0 ? -1.0 : cosine);
    const double force = -ka * (acos(cosine) - aHOH0);
    const V3 Fa = scale(pa, force / nba), Fc = scale(pc, force / nbc);
    const V3 Fb{-(Fa.x + Fc.x), -(Fa.y + Fc.y), -(Fa.z + Fc.z)};
    aA.x += Fa.x * imA; aA.y += Fa.y * imA; aA.z += Fa.z * imA;
    aC.x += Fc.x * imC; aC.y += Fc.y * imC; aC.z += Fc.z * imC;
    aO.x += Fb.x * imO; aO.y += Fb.y * imO; aO.z += Fb.z * imO;

    acc[o] += aO.x; acc[ld + o] += aO.y; acc[2 * ld + o] += aO.z;
    acc[h1] += aA.x; acc[ld + h1] += aA.y; acc[2 * ld + h1] += aA.z;
    acc[h2] += aC.x; acc[ld + h2] += aC.y; acc[2 * ld + h2] += aC.z;
}

// adds the bonded accelerations of the molecules whose oxygen column lies in [tgt_lo, tgt_hi)
int launch_spcfw_bonded(nbx_ctx *c, double *acc_out)
{
    const int mlo = (int)((c->tgt_lo + 2) / 3), mhi = (int)((c->tgt_hi + 2) / 3);
    if (mhi <= mlo) return NBX_OK;
    const int blocks = (mhi - mlo + 127) / 128;
    timer_begin(c, NBX_T_BONDED);
    spcfw_kernel<<<blocks, 128, 0, c->stream>>>(c->pos, c->npad, c->mass, mlo, mhi, c->rOH, c->aHOH, c->k_bond,
                                               c->k_angle, acc_out);
    timer_end(c, NBX_T_BONDED);
    NBX_CUDA(c, cudaGetLastError());
    return NBX_OK;
}

} // namespace nbx
