// ORACLE — TEST INFRASTRUCTURE ONLY (see yko_math.h header).
//
// CPU restatement of yuki's Material -> BSDF -> BxDF stack: materials/{matte,glass,metal,glossy}.rs and
// materials/bsdfs/{mod,lambertian,oren_nayar,specular,fresnel,microfacet,trowbridge_reitz}.rs.
// PARITY UNPINNED (no reference tests cover these); follows the cited lines, quirks included.
#pragma once
#include "yko_scene.h"

namespace yko {

// bsdfs/mod.rs:25-35
enum BxdfType : uint8_t {
    BXDF_NONE = 0,
    BXDF_REFLECTION = 1,
    BXDF_TRANSMISSION = 2,
    BXDF_DIFFUSE = 4,
    BXDF_GLOSSY = 8,
    BXDF_SPECULAR = 16,
    BXDF_ALL = 31,
};

struct BxdfSample {
    V3 wi{0, 0, 0};
    Spec f{0, 0, 0};
    float pdf = 0.0f;
    uint8_t sample_type = BXDF_NONE;
};

// bsdfs/mod.rs:225-300
inline float cos_theta(V3 w) { return w.z; }
inline float cos_2_theta(V3 w) { return w.z * w.z; }
inline float sin_2_theta(V3 w) { return fmax_(1.0f - cos_2_theta(w), 0.0f); }
inline float sin_theta(V3 w) { return std::sqrt(sin_2_theta(w)); }
inline float tan_theta(V3 w) { return sin_theta(w) / cos_theta(w); }
inline float tan_2_theta(V3 w) { return sin_2_theta(w) / cos_2_theta(w); }
inline float sin_phi(V3 w) {
    float st = sin_theta(w);
    return st == 0.0f ? 1.0f : clampf(w.y / st, -1.0f, 1.0f);  // 1.0 (not 0.0) when sin(theta)==0: reference quirk
}
inline float cos_phi(V3 w) {
    float st = sin_theta(w);
    return st == 0.0f ? 1.0f : clampf(w.x / st, -1.0f, 1.0f);
}
inline float sin_2_phi(V3 w) { return sin_phi(w) * sin_phi(w); }
inline float cos_2_phi(V3 w) { return cos_phi(w) * cos_phi(w); }
inline bool same_hemisphere(V3 w, V3 wp) { return w.z * wp.z > 0.0f; }
inline V3 spherical_direction(float sin_t, float cos_t, float phi) {
    return {sin_t * std::cos(phi), sin_t * std::sin(phi), cos_t};
}
inline bool refract(V3 wi, V3 n, float eta, V3* wt) {
    float cos_theta_i = dot_nv(n, wi);
    float sin_2_theta_i = fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f);
    float sin_2_theta_t = eta * eta * sin_2_theta_i;
    if (sin_2_theta_t >= 1.0f) return false;
    float cos_theta_t = std::sqrt(1.0f - sin_2_theta_t);
    *wt = (-wi) * eta + n * (eta * cos_theta_i - cos_theta_t);
    return true;
}
inline V3 reflect(V3 wo, V3 n) { return (-wo) + n * 2.0f * dot(wo, n); }

// bsdfs/fresnel.rs
inline Spec fresnel_dielectric(float eta_i_in, float eta_t_in, float cos_theta_i) {  // :21-51
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    bool entering = cos_theta_i > 0.0f;
    float eta_i = entering ? eta_i_in : eta_t_in;
    float eta_t = entering ? eta_t_in : eta_i_in;
    if (!entering) cos_theta_i = std::fabs(cos_theta_i);
    float sin_theta_i = std::sqrt(fmax_(1.0f - cos_theta_i * cos_theta_i, 0.0f));
    float sin_theta_t = eta_i / eta_t * sin_theta_i;
    if (sin_theta_t >= 1.0f) return spec1(1.0f);
    float cos_theta_t = std::sqrt(fmax_(1.0f - sin_theta_t * sin_theta_t, 0.0f));
    float r_par = ((eta_t * cos_theta_i) - (eta_i * cos_theta_t)) / ((eta_t * cos_theta_i) + (eta_i * cos_theta_t));
    float r_perp = ((eta_i * cos_theta_i) - (eta_t * cos_theta_t)) / ((eta_i * cos_theta_i) + (eta_t * cos_theta_t));
    return spec1(1.0f) * (r_par * r_par + r_perp * r_perp) / 2.0f;
}
inline Spec ssqrt(Spec v) { return {std::sqrt(v.r), std::sqrt(v.g), std::sqrt(v.b)}; }
inline Spec fresnel_conductor(Spec eta_i, Spec eta_t, Spec k, float cos_theta_i) {  // :68-96
    cos_theta_i = fmin_(std::fabs(cos_theta_i), 1.0f);
    Spec eta = eta_t / eta_i;
    Spec eta_k = k / eta_i;
    float c2 = cos_theta_i * cos_theta_i;
    float s2 = 1.0f - c2;
    Spec eta_2 = eta * eta;
    Spec eta_k_2 = eta_k * eta_k;
    Spec t0 = eta_2 - eta_k_2 - spec1(s2);                 // SubScalar: component - scalar
    Spec a2b2 = ssqrt(t0 * t0 + eta_2 * eta_k_2 * 4.0f);
    Spec t1 = a2b2 + spec1(c2);
    Spec a = ssqrt((a2b2 + t0) * 0.5f);
    Spec t2 = a * cos_theta_i * 2.0f;
    Spec rs = (t1 - t2) / (t1 + t2);
    Spec t3 = a2b2 * c2 + spec1(s2 * s2);
    Spec t4 = t2 * s2;
    Spec rp = rs * (t3 - t4) / (t3 + t4);
    return (rp + rs) * 0.5f;
}
inline Spec fresnel_schlick(Spec rs, float cos_theta_i) {  // :108-117
    cos_theta_i = clampf(cos_theta_i, -1.0f, 1.0f);
    float v = 1.0f - cos_theta_i;
    float p5 = (v * v) * (v * v) * v;
    return rs + (spec1(1.0f) - rs) * p5;
}

// bsdfs/trowbridge_reitz.rs
inline float tr_roughness_to_alpha(float roughness) {  // :22-30
    float x = std::log(fmax_(roughness, 0.001f));
    return 1.62142f + 0.819955f * x + 0.1734f * x * x + 0.0171201f * x * x * x + 0.000640711f * x * x * x * x;
}
inline float tr_d(float alpha, V3 wh) {  // :34-44
    float t2 = tan_2_theta(wh);
    if (std::isinf(t2)) return 0.0f;
    float alpha_2 = alpha * alpha;
    float cos_4 = cos_2_theta(wh) * cos_2_theta(wh);
    float e = (cos_2_phi(wh) / alpha_2 + sin_2_phi(wh) / alpha_2) * t2;
    return 1.0f / (PI_F * alpha_2 * cos_4 * (1.0f + e) * (1.0f + e));
}
inline float tr_lambda(float alpha_in, V3 w) {  // :46-58
    float abs_tan = std::fabs(tan_theta(w));
    if (std::isinf(abs_tan)) return 0.0f;
    float alpha = std::sqrt(cos_2_phi(w) * alpha_in * alpha_in + sin_2_phi(w) * alpha_in * alpha_in);
    float a2t2 = (alpha * abs_tan) * (alpha * abs_tan);
    return (-1.0f + std::sqrt(1.0f + a2t2)) / 2.0f;
}
inline float tr_g(float alpha, V3 wo, V3 wi) { return 1.0f / (1.0f + tr_lambda(alpha, wo) + tr_lambda(alpha, wi)); }
inline V3 tr_sample_wh(float alpha, V3 wo, V2 u) {  // :60-74 — full-distribution sampling
    float tan_theta_2 = alpha * alpha * u.x / (1.0f - u.x);
    float cos_t = 1.0f / std::sqrt(1.0f + tan_theta_2);
    float phi = 2.0f * PI_F * u.y;
    float sin_t = std::sqrt(fmax_(1.0f - cos_t * cos_t, 0.0f));
    V3 wh = spherical_direction(sin_t, cos_t, phi);
    return same_hemisphere(wo, wh) ? wh : -wh;
}
inline float tr_pdf(float alpha, V3 wh) { return tr_d(alpha, wh) * cos_theta(wh); }  // :76-78 (no abs)

enum LobeKind : uint8_t { LOBE_LAMBERT, LOBE_OREN_NAYAR, LOBE_SPEC_REFL, LOBE_SPEC_TRANS, LOBE_MICROFACET };
enum FresnelKind : uint8_t { FRESNEL_DIELECTRIC, FRESNEL_CONDUCTOR, FRESNEL_SCHLICK };

struct Lobe {
    LobeKind kind;
    Spec r;                   // reflectance / R / T
    float a = 0, b = 0;       // Oren-Nayar
    float eta_i = 1, eta_t = 1;  // dielectric (reflection's Fresnel and transmission)
    FresnelKind fresnel = FRESNEL_DIELECTRIC;
    Spec c_eta_i{1, 1, 1}, c_eta_t{1, 1, 1}, c_k{0, 0, 0};  // conductor
    Spec rs{0, 0, 0};         // schlick
    float alpha = 0;          // Trowbridge-Reitz

    uint8_t flags() const {
        switch (kind) {
            case LOBE_LAMBERT:
            case LOBE_OREN_NAYAR: return BXDF_DIFFUSE | BXDF_REFLECTION;
            case LOBE_SPEC_REFL: return BXDF_SPECULAR | BXDF_REFLECTION;
            case LOBE_SPEC_TRANS: return BXDF_SPECULAR | BXDF_TRANSMISSION;
            default: return BXDF_REFLECTION | BXDF_GLOSSY;
        }
    }
    bool matches(uint8_t t) const { return (t & flags()) == flags(); }  // t.contains(self.flags())

    Spec eval_fresnel(float c) const {
        switch (fresnel) {
            case FRESNEL_DIELECTRIC: return fresnel_dielectric(eta_i, eta_t, c);
            case FRESNEL_CONDUCTOR: return fresnel_conductor(c_eta_i, c_eta_t, c_k, c);
            default: return fresnel_schlick(rs, c);
        }
    }

    // Bxdf::f(wo, wi)
    Spec f(V3 wo, V3 wi) const {
        switch (kind) {
            case LOBE_LAMBERT: return r * FRAC_1_PI_F;  // lambertian.rs:21-23
            case LOBE_OREN_NAYAR: {
                // oren_nayar.rs:29-53 — the impl names its parameters (wi, wo), i.e. swapped relative to the
                // trait's (wo, wi). Restated as written: `pi` is the first argument.
                V3 pi = wo, po = wi;
                float sin_i = sin_theta(pi), sin_o = sin_theta(po);
                float max_cos = 0.0f;
                if (sin_i > 1e-4f && sin_o > 1e-4f) {
                    float sin_phi_i = sin_phi(pi), cos_phi_i = cos_phi(pi);
                    float sin_phi_o = sin_phi(po), cos_phi_o = cos_phi(po);
                    float d_cos = cos_phi_i * cos_phi_o + sin_phi_i * sin_phi_o;
                    max_cos = fmax_(d_cos, 0.0f);
                }
                float sin_alpha, tan_beta;
                if (std::fabs(cos_theta(pi)) > std::fabs(cos_theta(po))) {
                    sin_alpha = sin_o;
                    tan_beta = sin_i / std::fabs(cos_theta(pi));
                } else {
                    sin_alpha = sin_i;
                    tan_beta = sin_o / std::fabs(cos_theta(po));
                }
                return r * FRAC_1_PI_F * (a + b * max_cos * sin_alpha * tan_beta);
            }
            case LOBE_SPEC_REFL:
            case LOBE_SPEC_TRANS: return spec1(0.0f);
            case LOBE_MICROFACET: {  // microfacet.rs:53-74
                float cos_o = std::fabs(cos_theta(wo)), cos_i = std::fabs(cos_theta(wi));
                if (cos_i == 0.0f || cos_o == 0.0f) return spec1(0.0f);
                V3 wh = wi + wo;
                if (wh.x == 0.0f && wh.y == 0.0f && wh.z == 0.0f) return spec1(0.0f);
                wh = normalized(wh);
                Spec fr = eval_fresnel(dot(wi, faceforward(wh, v3(0.0f, 0.0f, 1.0f))));
                return r * tr_d(alpha, wh) * tr_g(alpha, wo, wi) * fr / (4.0f * cos_i * cos_o);
            }
        }
        return spec1(0.0f);
    }
    float pdf(V3 wo, V3 wi) const {
        switch (kind) {
            case LOBE_LAMBERT:
            case LOBE_OREN_NAYAR: return same_hemisphere(wo, wi) ? std::fabs(cos_theta(wi)) * FRAC_1_PI_F : 0.0f;
            case LOBE_SPEC_REFL:
            case LOBE_SPEC_TRANS: return 1.0f;
            case LOBE_MICROFACET: {  // microfacet.rs:101-108
                if (!same_hemisphere(wo, wi)) return 0.0f;
                V3 wh = normalized(wo + wi);
                return tr_pdf(alpha, wh) / (4.0f * dot(wo, wh));
            }
        }
        return 0.0f;
    }
    BxdfSample sample_f(V3 wo, V2 u) const {
        BxdfSample s;
        switch (kind) {
            case LOBE_LAMBERT:
            case LOBE_OREN_NAYAR: {  // lambertian.rs:25-40, oren_nayar.rs:55-70
                V3 wi = cosine_sample_hemisphere(u);
                if (wo.z < 0.0f) wi.z *= -1.0f;
                s.wi = wi; s.pdf = pdf(wo, wi); s.f = f(wo, wi); s.sample_type = flags();
            } break;
            case LOBE_SPEC_REFL: {  // specular.rs:26-37
                V3 wi = v3(-wo.x, -wo.y, wo.z);
                s.wi = wi;
                s.f = r * eval_fresnel(cos_theta(wi)) / std::fabs(cos_theta(wi));
                s.pdf = 1.0f; s.sample_type = flags();
            } break;
            case LOBE_SPEC_TRANS: {  // specular.rs:69-92 — no (eta_i/eta_t)^2 radiance scaling
                bool entering = cos_theta(wo) > 0.0f;
                float ei = entering ? eta_i : eta_t, et = entering ? eta_t : eta_i;
                V3 wi;
                if (!refract(wo, faceforward(v3(0.0f, 0.0f, 1.0f), wo), ei / et, &wi)) return s;
                s.wi = wi;
                s.f = r * (spec1(1.0f) - fresnel_dielectric(eta_i, eta_t, cos_theta(wi))) / std::fabs(cos_theta(wi));
                s.pdf = 1.0f; s.sample_type = flags();
            } break;
            case LOBE_MICROFACET: {  // microfacet.rs:76-99
                if (wo.z == 0.0f) return s;
                V3 wh = tr_sample_wh(alpha, wo, u);
                if (dot(wo, wh) < 0.0f) return s;
                V3 wi = reflect(wo, wh);
                if (!same_hemisphere(wo, wi)) return s;
                s.wi = wi;
                s.pdf = tr_pdf(alpha, wh) / (4.0f * dot(wo, wh));
                s.f = f(wo, wi);
                s.sample_type = flags();
            } break;
        }
        return s;
    }
};

// bsdfs/mod.rs:75-223
struct Bsdf {
    Lobe lobes[2];
    int n_lobes = 0;
    V3 n_geom, n_shading, s_shading, t_shading;

    explicit Bsdf(const SurfaceInteraction& si) {  // :87-99
        n_geom = si.n;
        n_shading = si.sh_n;
        s_shading = normalized(si.sh_dpdu);
        t_shading = cross(n_shading, s_shading);
    }
    void add(const Lobe& l) { lobes[n_lobes++] = l; }
    V3 world_to_local(V3 v) const { return {dot(v, s_shading), dot(v, t_shading), dot_nv(v, n_shading)}; }
    V3 local_to_world(V3 v) const {
        return {s_shading.x * v.x + t_shading.x * v.y + n_shading.x * v.z,
                s_shading.y * v.x + t_shading.y * v.y + n_shading.y * v.z,
                s_shading.z * v.x + t_shading.z * v.y + n_shading.z * v.z};
    }
    Spec f(V3 wo_world, V3 wi_world, uint8_t type) const {  // :125-147
        V3 wo = world_to_local(wo_world), wi = world_to_local(wi_world);
        bool refl = dot_nv(wi_world, n_geom) * dot_nv(wo_world, n_geom) > 0.0f;
        Spec f = spec1(0.0f);
        for (int i = 0; i < n_lobes; ++i) {
            const Lobe& b = lobes[i];
            if (b.matches(type) && ((refl && (b.flags() & BXDF_REFLECTION)) || (!refl && (b.flags() & BXDF_TRANSMISSION))))
                f += b.f(wo, wi);
        }
        return f;
    }
    BxdfSample sample_f(V3 wo_world, V2 u, uint8_t type) const {  // :150-222
        int matching = 0;
        for (int i = 0; i < n_lobes; ++i) matching += lobes[i].matches(type) ? 1 : 0;
        if (matching == 0) return BxdfSample{};
        float fl = std::floor(u.x * (float)matching);
        int comp = std::min(fl > 0.0f ? (int)fl : 0, matching - 1);
        const Lobe* bxdf = nullptr;
        for (int i = 0, k = 0; i < n_lobes; ++i)
            if (lobes[i].matches(type) && k++ == comp) { bxdf = &lobes[i]; break; }
        V3 wo = world_to_local(wo_world);
        V2 u_remapped{u.x * (float)(matching - comp), u.y};  // :176 reference quirk, reproduced
        BxdfSample s = bxdf->sample_f(wo, u_remapped);
        if (s.pdf == 0.0f) return BxdfSample{};
        V3 wi_local = s.wi;
        V3 wi_world = local_to_world(wi_local);
        bool specular = (bxdf->flags() & BXDF_SPECULAR) != 0;
        if (!specular && matching > 1)
            for (int i = 0; i < n_lobes; ++i)
                if (&lobes[i] != bxdf && lobes[i].matches(type)) s.pdf += lobes[i].pdf(wo, wi_local);
        if (matching > 1) s.pdf /= (float)matching;
        if (!specular && matching > 1) {
            bool refl = dot_nv(wi_world, n_geom) * dot_nv(wo_world, n_geom) > 0.0f;
            s.f = spec1(0.0f);
            for (int i = 0; i < n_lobes; ++i) {
                const Lobe& b = lobes[i];
                if (b.matches(type) && ((refl && (b.flags() & BXDF_REFLECTION)) || (!refl && (b.flags() & BXDF_TRANSMISSION))))
                    s.f += b.f(wo, wi_local);
            }
        }
        s.wi = wi_world;
        return s;
    }
};

// materials/{matte,glass,metal,glossy}.rs compute_scattering_functions
inline Bsdf compute_scattering_functions(const Scene& scene, const Material& m, const SurfaceInteraction& si) {
    Bsdf bsdf(si);
    auto tex = [&](int i) { return texture_eval(scene.textures[m.tex[i]], si); };
    switch (m.kind) {
        case MAT_MATTE: {  // matte.rs:22-40
            Spec kd = tex(0);
            float sigma = tex(1).r;
            if (!is_black(kd)) {
                Lobe l{};
                l.r = kd;
                if (sigma == 0.0f) l.kind = LOBE_LAMBERT;
                else {  // oren_nayar.rs:18-25 — sigma in RADIANS
                    l.kind = LOBE_OREN_NAYAR;
                    float s2 = sigma * sigma;
                    l.a = 1.0f - (s2 / (2.0f * (s2 + 0.33f)));
                    l.b = 0.45f * s2 / (s2 + 0.09f);
                }
                bsdf.add(l);
            }
        } break;
        case MAT_GLASS: {  // glass.rs:27-45
            Lobe r{};
            r.kind = LOBE_SPEC_REFL; r.r = tex(0); r.fresnel = FRESNEL_DIELECTRIC; r.eta_i = 1.0f; r.eta_t = m.eta;
            bsdf.add(r);
            Lobe t{};
            t.kind = LOBE_SPEC_TRANS; t.r = tex(1); t.eta_i = 1.0f; t.eta_t = m.eta;
            bsdf.add(t);
        } break;
        case MAT_METAL: {  // metal.rs:34-61
            float rough = tex(2).r;
            if (m.remap_roughness) rough = tr_roughness_to_alpha(rough);
            Lobe l{};
            l.kind = LOBE_MICROFACET; l.r = spec1(1.0f); l.fresnel = FRESNEL_CONDUCTOR;
            l.c_eta_i = spec1(1.0f); l.c_eta_t = tex(0); l.c_k = tex(1);
            l.alpha = fmax_(rough, 0.001f);  // trowbridge_reitz.rs:16-20
            bsdf.add(l);
        } break;
        case MAT_GLOSSY: {  // glossy.rs:32-58
            float rough = tex(1).r;
            if (m.remap_roughness) rough = tr_roughness_to_alpha(rough);
            Lobe l{};
            l.kind = LOBE_MICROFACET; l.r = spec1(1.0f); l.fresnel = FRESNEL_SCHLICK; l.rs = tex(0);
            l.alpha = fmax_(rough * rough, 0.001f);  // glossy.rs:49 squared roughness
            bsdf.add(l);
        } break;
    }
    return bsdf;
}

}  // namespace yko
