// scene_text.cpp — a line-oriented scene description that drives the same builder calls as the C ABI.
//
// The reference has no scene files: its ten scenes are C++ functions (mort.cu:129-631) and adding one means
// recompiling.  Here every statement is exactly one mort_add_* call (world::add overloads, world.cuh:27-90), in
// file order, so array slots — and with them primitive ids, list orders and the BVH the reference would have
// built — are decided by the file alone.  The Scene keeps a journal of the builder calls it received
// (scene.cpp), so any scene built through the API, including the ten shipped ones, can be written back out
// and replays to the same arrays bit for bit.
//
//   # comment                         blank lines and everything after '#' are ignored
//   [NAME =] STATEMENT                NAME becomes an alias for the handle the statement returns
//
//   solid R G B                       checker SCALE EVEN ODD          image FILE.ppm | image @K
//   noise SCALE [at K]                K = position in the unseeded host rand() stream the tables are drawn from
//                                     (without it they come from the context's stream, like mort_add_noise)
//   lambertian TEX | R G B            metal R G B FUZZ                dielectric IOR
//   light TEX | R G B                 isotropic TEX | R G B
//   sphere CX CY CZ R MAT [hidden]    moving_sphere C0(3) C1(3) R MAT [hidden]
//   quad Q(3) U(3) V(3) MAT [hidden]  translate OBJ X Y Z [hidden]    rotate_y OBJ DEGREES [hidden]
//   medium BOUNDARY DENSITY MAT [hidden]
//   list [hidden]                     add LIST OBJ                    bvh LIST [hidden]
//   box A(3) B(3) MAT                 rotated_box SIZE(3) TRANSLATION(3) DEGREES MAT        (utils.h:51-96)
//   host_rand_skip N                  advance the context's host rand() stream (scene code that calls rand())
//   camera width N | aspect A | spp N | depth N | vfov N | background R G B | lookfrom X Y Z | lookat X Y Z |
//          vup X Y Z | defocus_angle A | focus_dist D | light OBJ|none
//
// `hidden` = reachable only through a parent (the reference's `skip`).  Handles are aliases or the canonical
// kind+slot spelling the journal writes: sph qd tr ry med list bvh / lam met die lgt iso / sol chk img noi.
#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "scene.hpp"

namespace mort {
namespace {

struct Bound { Handle h; char cat; };

struct Parser {
    Scene& s; HostRng& rng; std::string asset_dir, path, err;
    std::map<std::string, Bound> alias;
    std::vector<std::string> tok; size_t at = 0; int line_no = 0;

    bool fail(const std::string& m) { err = path + ":" + std::to_string(line_no) + ": " + m; return false; }
    bool more() const { return at < tok.size(); }
    bool num(float& v) {
        if (!more()) return fail("number expected");
        char* e = nullptr; errno = 0; v = strtof(tok[at].c_str(), &e);
        if (e == tok[at].c_str() || *e) return fail("number expected, got '" + tok[at] + "'");
        at++; return true;
    }
    bool integer(long long& v) {
        if (!more()) return fail("integer expected");
        char* e = nullptr; v = strtoll(tok[at].c_str(), &e, 10);
        if (e == tok[at].c_str() || *e) return fail("integer expected, got '" + tok[at] + "'");
        at++; return true;
    }
    bool v3(V3& v) { return num(v.x) && num(v.y) && num(v.z); }
    bool is_number(const std::string& t) const { char* e = nullptr; strtof(t.c_str(), &e); return e != t.c_str() && !*e; }
    bool handle(char cat, Handle& h) {
        if (!more()) return fail("handle expected");
        const std::string t = tok[at++];
        auto it = alias.find(t);
        if (it != alias.end()) { if (it->second.cat != cat) return fail("'" + t + "' is not a " + (cat == 'o' ? "hittable" : cat == 'm' ? "material" : "texture")); h = it->second.h; return true; }
        static const struct { const char* p; char cat; int type; } kinds[] = {
            {"sph", 'o', MORT_OBJ_SPHERE}, {"qd", 'o', MORT_OBJ_QUAD}, {"tr", 'o', MORT_OBJ_TRANSLATE}, {"ry", 'o', MORT_OBJ_ROTATE_Y},
            {"med", 'o', MORT_OBJ_CONSTANT_MEDIUM}, {"list", 'o', MORT_OBJ_HITTABLE_LIST}, {"bvh", 'o', MORT_OBJ_BVH},
            {"lam", 'm', MORT_MAT_LAMBERTIAN}, {"met", 'm', MORT_MAT_METAL}, {"die", 'm', MORT_MAT_DIELECTRIC}, {"lgt", 'm', MORT_MAT_DIFFUSE_LIGHT},
            {"iso", 'm', MORT_MAT_ISOTROPIC}, {"sol", 't', MORT_TEX_SOLID}, {"chk", 't', MORT_TEX_CHECKER}, {"img", 't', MORT_TEX_IMAGE}, {"noi", 't', MORT_TEX_NOISE}};
        for (const auto& k : kinds) {
            const size_t n = strlen(k.p);
            if (k.cat == cat && t.compare(0, n, k.p) == 0 && t.size() > n && t.find_first_not_of("0123456789", n) == std::string::npos) {
                h = Handle{k.type, atoi(t.c_str() + n)};
                if (!exists(h, cat)) return fail("'" + t + "' does not exist yet");
                return true;
            }
        }
        return fail("unknown handle '" + t + "'");
    }
    bool exists(Handle h, char cat) const {
        size_t n = 0;
        if (cat == 'o') switch (h.type) { case MORT_OBJ_SPHERE: n = s.spheres.size(); break; case MORT_OBJ_QUAD: n = s.quads.size(); break; case MORT_OBJ_TRANSLATE: n = s.translates.size(); break;
            case MORT_OBJ_ROTATE_Y: n = s.rotates.size(); break; case MORT_OBJ_CONSTANT_MEDIUM: n = s.media.size(); break; case MORT_OBJ_HITTABLE_LIST: n = s.lists.size(); break; case MORT_OBJ_BVH: n = s.bvhs.size(); break; }
        else if (cat == 'm') switch (h.type) { case MORT_MAT_LAMBERTIAN: n = s.lambertians.size(); break; case MORT_MAT_METAL: n = s.metals.size(); break; case MORT_MAT_DIELECTRIC: n = s.dielectrics.size(); break;
            case MORT_MAT_DIFFUSE_LIGHT: n = s.lights.size(); break; case MORT_MAT_ISOTROPIC: n = s.isotropics.size(); break; }
        else switch (h.type) { case MORT_TEX_SOLID: n = s.solids.size(); break; case MORT_TEX_CHECKER: n = s.checkers.size(); break; case MORT_TEX_IMAGE: n = s.images.size(); break; case MORT_TEX_NOISE: n = s.noises.size(); break; }
        return h.idx >= 0 && (size_t)h.idx < n;
    }
    bool hidden(bool& skip) {
        skip = false;
        if (more() && tok[at] == "hidden") { skip = true; at++; }
        return true;
    }
    // TEX | R G B
    bool tex_or_colour(Handle& t) {
        if (at + 2 < tok.size() && is_number(tok[at])) { V3 c; if (!v3(c)) return false; t = s.add_solid(c); return true; }
        return handle('t', t);
    }

    bool statement(Bound& out, bool& has_result) {
        has_result = true;
        const std::string k = tok[at++];
        Handle a, b, m; V3 p, q, r; float x, y; bool skip; long long n;
        if (k == "solid") { if (!v3(p)) return false; out = {s.add_solid(p), 't'}; }
        else if (k == "checker") { if (!num(x) || !handle('t', a) || !handle('t', b)) return false; out = {s.add_checker(x, a, b), 't'}; }
        else if (k == "image") {
            if (!more()) return fail("image: file name expected");
            const std::string name = tok[at++];
            ImageRec im; bool ok;
            if (name[0] == '@') ok = load_ppm(path + ".img" + name.substr(1) + ".ppm", im);
            else { ok = load_ppm(asset_dir + "/" + name, im); if (!ok) { const size_t sl = path.find_last_of('/'); ok = load_ppm((sl == std::string::npos ? std::string(".") : path.substr(0, sl)) + "/" + name, im); } }
            if (!ok) return fail("cannot read image '" + name + "' (binary PPM, P6 / 255)");
            out = {s.add_image(im.rgb.data(), im.width, im.height, name[0] == '@' ? nullptr : name.c_str()), 't'};
        }
        else if (k == "noise") {
            if (!num(x)) return false;
            if (more() && tok[at] == "at") { at++; if (!integer(n) || n < 0) return fail("noise: stream position expected after 'at'"); HostRng g(1); g.skip((uint64_t)n); out = {s.add_noise(x, g), 't'}; }
            else out = {s.add_noise(x, rng), 't'};
        }
        else if (k == "lambertian") { if (!tex_or_colour(a)) return false; out = {s.add_lambertian(a), 'm'}; }
        else if (k == "metal") { if (!v3(p) || !num(x)) return false; out = {s.add_metal(p, x), 'm'}; }
        else if (k == "dielectric") { if (!num(x)) return false; out = {s.add_dielectric(x), 'm'}; }
        else if (k == "light") { if (!tex_or_colour(a)) return false; out = {s.add_diffuse_light(a), 'm'}; }
        else if (k == "isotropic") { if (!tex_or_colour(a)) return false; out = {s.add_isotropic(a), 'm'}; }
        else if (k == "sphere") { if (!v3(p) || !num(x) || !handle('m', m) || !hidden(skip)) return false; out = {s.add_sphere(p, x, m, skip), 'o'}; }
        else if (k == "moving_sphere") { if (!v3(p) || !v3(q) || !num(x) || !handle('m', m) || !hidden(skip)) return false; out = {s.add_moving_sphere(p, q, x, m, skip), 'o'}; }
        else if (k == "quad") { if (!v3(p) || !v3(q) || !v3(r) || !handle('m', m) || !hidden(skip)) return false; out = {s.add_quad(p, q, r, m, skip), 'o'}; }
        else if (k == "translate") { if (!handle('o', a) || !v3(p) || !hidden(skip)) return false; out = {s.add_translate(a, p, skip), 'o'}; }
        else if (k == "rotate_y") { if (!handle('o', a) || !num(x) || !hidden(skip)) return false; out = {s.add_rotate_y(a, x, skip), 'o'}; }
        else if (k == "medium") { if (!handle('o', a) || !num(x) || !handle('m', m) || !hidden(skip)) return false; if (!(x > 0)) return fail("medium: density must be positive"); out = {s.add_constant_medium(a, x, m, skip), 'o'}; }
        else if (k == "list") { if (!hidden(skip)) return false; out = {s.add_list(skip), 'o'}; }
        else if (k == "add") { if (!handle('o', a) || !handle('o', b)) return false; if (s.list_add(a, b) != 0) return fail("add: first handle is not a list"); has_result = false; }
        else if (k == "bvh") { if (!handle('o', a) || !hidden(skip)) return false; if (a.type != MORT_OBJ_HITTABLE_LIST) return fail("bvh: handle is not a list"); out = {s.add_bvh(a, skip), 'o'}; }
        else if (k == "box") { if (!v3(p) || !v3(q) || !handle('m', m)) return false; s.box(p, q, m); has_result = false; }
        else if (k == "rotated_box") { if (!v3(p) || !v3(q) || !num(x) || !handle('m', m)) return false; out = {s.rotated_box(p, q, x, m), 'o'}; }
        else if (k == "host_rand_skip") { if (!integer(n) || n < 0) return fail("host_rand_skip: count expected"); rng.skip((uint64_t)n); has_result = false; }
        else if (k == "camera") { has_result = false; return camera(); }
        else return fail("unknown statement '" + k + "'");
        (void)y;
        if (more()) return fail("unexpected '" + tok[at] + "'");
        return true;
    }
    bool camera() {
        Camera& c = s.cam; long long n; float x;
        while (more()) {
            const std::string k = tok[at++];
            if (k == "width") { if (!integer(n) || n < 1) return fail("camera width: positive integer expected"); c.image_width = (int)n; }
            else if (k == "aspect") { if (!num(x) || !(x > 0)) return fail("camera aspect: positive number expected"); c.aspect_ratio = x; }
            else if (k == "spp") { if (!integer(n) || n < 1) return fail("camera spp: positive integer expected"); c.samples_per_pixel = (int)n; }
            else if (k == "depth") { if (!integer(n) || n < 0) return fail("camera depth: integer >= 0 expected"); c.bounce_limit = (int)n; }
            else if (k == "vfov") { if (!integer(n)) return false; c.vfov = (int)n; }
            else if (k == "background") { if (!v3(c.background)) return false; }
            else if (k == "lookfrom") { if (!v3(c.lookfrom)) return false; }
            else if (k == "lookat") { if (!v3(c.lookat)) return false; }
            else if (k == "vup") { if (!v3(c.vup)) return false; }
            else if (k == "defocus_angle") { if (!num(c.defocus_angle)) return false; }
            else if (k == "focus_dist") { if (!num(c.focus_dist)) return false; }
            else if (k == "light") {
                if (more() && tok[at] == "none") { at++; c.light_obj_type = -1; c.light_obj_idx = 0; }
                else { Handle h; if (!handle('o', h)) return false; c.light_obj_type = h.type; c.light_obj_idx = h.idx; }
            }
            else return fail("unknown camera field '" + k + "'");
        }
        return true;
    }
};

}  // namespace

bool load_scene_text(Scene& s, HostRng& rng, const std::string& path, const std::string& asset_dir, std::string* err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { if (err) *err = "cannot open " + path; return false; }
    std::string text; char buf[4096]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) text.append(buf, n);
    fclose(f);
    s.clear();
    Parser P{s, rng, asset_dir, path, std::string(), {}, {}, 0, 0};
    std::istringstream in(text); std::string line;
    while (std::getline(in, line)) {
        P.line_no++;
        const size_t hash = line.find('#');
        if (hash != std::string::npos) line.erase(hash);
        std::istringstream ls(line); std::string t; P.tok.clear(); P.at = 0;
        while (ls >> t) P.tok.push_back(t);
        if (P.tok.empty()) continue;
        std::string name;
        if (P.tok.size() >= 2 && P.tok[1] == "=") { name = P.tok[0]; P.at = 2; if (P.tok.size() < 3) { P.fail("statement expected after '='"); break; } }
        Bound b{Handle{-1, -1}, 'o'}; bool has = false;
        if (!P.statement(b, has)) break;
        if (!name.empty()) {
            if (!has) { P.fail("this statement returns no handle to name"); break; }
            P.alias[name] = b;
        }
    }
    if (!P.err.empty()) { if (err) *err = P.err; s.clear(); return false; }
    s.cam.initialize();
    return true;
}

bool dump_scene_text(const Scene& s, const std::string& path, std::string* err) {
    if (!s.journal_complete) { if (err) *err = "this scene was not built through builder calls (binary dump or explicit noise tables): it has no text form"; return false; }
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { if (err) *err = "cannot write " + path; return false; }
    fprintf(f, "# mort-b200 scene text: one builder call per statement, in call order (grammar: mort_b200/csrc/scene_text.cpp)\n");
    for (const std::string& l : s.journal) fprintf(f, "%s\n", l.c_str());
    const Camera& c = s.cam;
    fprintf(f, "camera width %d aspect %.9g spp %d depth %d vfov %d\n", c.image_width, (double)c.aspect_ratio, c.samples_per_pixel, c.bounce_limit, c.vfov);
    fprintf(f, "camera background %.9g %.9g %.9g\n", (double)c.background.x, (double)c.background.y, (double)c.background.z);
    fprintf(f, "camera lookfrom %.9g %.9g %.9g lookat %.9g %.9g %.9g vup %.9g %.9g %.9g\n", (double)c.lookfrom.x, (double)c.lookfrom.y, (double)c.lookfrom.z,
            (double)c.lookat.x, (double)c.lookat.y, (double)c.lookat.z, (double)c.vup.x, (double)c.vup.y, (double)c.vup.z);
    fprintf(f, "camera defocus_angle %.9g focus_dist %.9g\n", (double)c.defocus_angle, (double)c.focus_dist);
    // the reference leaves a stale or out-of-range handle in some scenes (App. A): keep it verbatim when it has a spelling
    const std::string lt = c.light_obj_type < 0 ? std::string("none") : handle_token(Handle{c.light_obj_type, c.light_obj_idx}, 'o');
    fprintf(f, "camera light %s\n", lt.c_str());
    bool ok = fclose(f) == 0;
    for (size_t k = 0; ok && k < s.images.size(); k++) {
        const ImageRec& im = s.images[k];
        if (!im.source.empty() || im.rgb.empty()) continue;
        const std::string ip = path + ".img" + std::to_string(k) + ".ppm";
        FILE* g = fopen(ip.c_str(), "wb");
        if (!g) { ok = false; break; }
        fprintf(g, "P6\n%d %d\n255\n", im.width, im.height);
        ok = fwrite(im.rgb.data(), 1, im.rgb.size(), g) == im.rgb.size();
        ok = (fclose(g) == 0) && ok;
    }
    if (!ok && err) *err = "cannot write " + path;
    return ok;
}

}  // namespace mort
