// dvren_render -- JSON scene description -> one rendered frame as a binary PPM, on the B200 runtime.
//
// Caller-side companion of the hot path ("next" row 3).  Behavioural contract taken from
// the reference tool (apps/dvren_render/main.cpp:77-370): same configuration keys and defaults, same
// exit codes, same stdout lines ("Forward stats: rays=.. samples=.. total_ms=..", "Workspace bytes ..",
// "Wrote <absolute path>"), same P6 quantisation round(clamp(v, 0, 1) * 255).  The JSON reader below is
// written for this tool (the reference vendors nlohmann/json); it accepts RFC 8259 documents.
//
//   dvren_render <config.json> [output.ppm]
//
//   { "render": { "width", "height", "t_far", "dt", "max_steps"            required
//                 "t_near" = 0, "seed" = 0, "sampling_mode" = "fixed" | "stratified",
//                 "roi": { "x" = 0, "y" = 0, "width" = W, "height" = H },
//                 "camera": { "model" = "pinhole" | "orthographic", "K": [9], "c2w": [12], "ortho_scale" = 1 },
//                 "options": { "use_fused_path" = true, "enable_graph" = false, "capture_stats" = true } },
//     "volume": { "size": [nx, ny, nz], "density": [nz*ny*nx], "color": [nz*ny*nx*3] (default: grey = density),
//                 "bbox_min": [3], "bbox_max": [3], "interp" = "linear" | "nearest", "oob" = "zero" | "clamp" },
//     "output": { "path" = "frame.ppm" } }
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <map>
#include <memory>
#include <optional>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "dvren/core/context.hpp"
#include "dvren/core/plan.hpp"
#include "dvren/fields/dense_grid.hpp"
#include "dvren/render/renderer.hpp"

namespace {

// ---------------------------------------------------------------------------------------------
// a small JSON document model + recursive-descent reader
// ---------------------------------------------------------------------------------------------
struct JsonError : std::runtime_error {
    using std::runtime_error::runtime_error;
};

class Json {
public:
    enum class Kind { kNull, kBool, kNumber, kString, kArray, kObject };

    Kind kind = Kind::kNull;
    bool boolean = false;
    double number = 0.0;
    std::string text;
    std::vector<Json> items;
    std::vector<std::pair<std::string, Json>> members;   // insertion order kept, later duplicates win on lookup

    [[nodiscard]] const Json* find(const std::string& key) const {
        if (kind != Kind::kObject) return nullptr;
        const Json* hit = nullptr;
        for (const auto& m : members)
            if (m.first == key) hit = &m.second;
        return hit;
    }
    [[nodiscard]] bool has(const std::string& key) const { return find(key) != nullptr; }
    [[nodiscard]] const Json& at(const std::string& key) const {
        const Json* j = find(key);
        if (j == nullptr) throw JsonError("key '" + key + "' not found");
        return *j;
    }
    [[nodiscard]] double as_number(const char* what) const {
        if (kind != Kind::kNumber) throw JsonError(std::string(what) + " must be a number");
        return number;
    }
    [[nodiscard]] const std::string& as_string(const char* what) const {
        if (kind != Kind::kString) throw JsonError(std::string(what) + " must be a string");
        return text;
    }
    [[nodiscard]] bool as_bool(const char* what) const {
        if (kind != Kind::kBool) throw JsonError(std::string(what) + " must be true or false");
        return boolean;
    }
};

class JsonReader {
public:
    explicit JsonReader(const std::string& source) : s_(source) {}

    Json parse_document() {
        Json v = parse_value(0);
        skip_space();
        if (pos_ != s_.size()) fail("trailing characters after the document");
        return v;
    }

private:
    const std::string& s_;
    size_t pos_ = 0;

    [[noreturn]] void fail(const std::string& what) const {
        size_t line = 1, col = 1;
        for (size_t i = 0; i < pos_ && i < s_.size(); ++i) {
            if (s_[i] == '\n') { ++line; col = 1; } else { ++col; }
        }
        throw JsonError("JSON parse error at line " + std::to_string(line) + ", column " + std::to_string(col) + ": " + what);
    }
    void skip_space() {
        while (pos_ < s_.size() && (s_[pos_] == ' ' || s_[pos_] == '\t' || s_[pos_] == '\n' || s_[pos_] == '\r')) ++pos_;
    }
    bool consume(char c) {
        skip_space();
        if (pos_ < s_.size() && s_[pos_] == c) { ++pos_; return true; }
        return false;
    }
    void expect_word(const char* word) {
        for (const char* p = word; *p; ++p, ++pos_)
            if (pos_ >= s_.size() || s_[pos_] != *p) fail(std::string("expected '") + word + "'");
    }

    Json parse_value(int depth) {
        if (depth > 64) fail("nesting too deep");
        skip_space();
        if (pos_ >= s_.size()) fail("unexpected end of input");
        Json v;
        const char c = s_[pos_];
        if (c == '{') {
            ++pos_;
            v.kind = Json::Kind::kObject;
            if (consume('}')) return v;
            do {
                skip_space();
                if (pos_ >= s_.size() || s_[pos_] != '"') fail("expected a member name");
                std::string key = parse_string();
                if (!consume(':')) fail("expected ':'");
                v.members.emplace_back(std::move(key), parse_value(depth + 1));
            } while (consume(','));
            if (!consume('}')) fail("expected ',' or '}'");
        } else if (c == '[') {
            ++pos_;
            v.kind = Json::Kind::kArray;
            if (consume(']')) return v;
            do {
                v.items.push_back(parse_value(depth + 1));
            } while (consume(','));
            if (!consume(']')) fail("expected ',' or ']'");
        } else if (c == '"') {
            v.kind = Json::Kind::kString;
            v.text = parse_string();
        } else if (c == 't') {
            expect_word("true");
            v.kind = Json::Kind::kBool;
            v.boolean = true;
        } else if (c == 'f') {
            expect_word("false");
            v.kind = Json::Kind::kBool;
        } else if (c == 'n') {
            expect_word("null");
        } else if (c == '-' || (c >= '0' && c <= '9')) {
            v.kind = Json::Kind::kNumber;
            v.number = parse_number();
        } else {
            fail("unexpected character");
        }
        return v;
    }

    double parse_number() {
        const size_t start = pos_;
        if (s_[pos_] == '-') ++pos_;
        if (pos_ >= s_.size() || s_[pos_] < '0' || s_[pos_] > '9') fail("malformed number");
        if (s_[pos_] == '0') {
            ++pos_;
        } else {
            while (pos_ < s_.size() && s_[pos_] >= '0' && s_[pos_] <= '9') ++pos_;
        }
        if (pos_ < s_.size() && s_[pos_] == '.') {
            ++pos_;
            if (pos_ >= s_.size() || s_[pos_] < '0' || s_[pos_] > '9') fail("malformed fraction");
            while (pos_ < s_.size() && s_[pos_] >= '0' && s_[pos_] <= '9') ++pos_;
        }
        if (pos_ < s_.size() && (s_[pos_] == 'e' || s_[pos_] == 'E')) {
            ++pos_;
            if (pos_ < s_.size() && (s_[pos_] == '+' || s_[pos_] == '-')) ++pos_;
            if (pos_ >= s_.size() || s_[pos_] < '0' || s_[pos_] > '9') fail("malformed exponent");
            while (pos_ < s_.size() && s_[pos_] >= '0' && s_[pos_] <= '9') ++pos_;
        }
        return std::strtod(s_.substr(start, pos_ - start).c_str(), nullptr);
    }

    static void append_utf8(std::string& out, uint32_t cp) {
        if (cp < 0x80) {
            out.push_back(static_cast<char>(cp));
        } else if (cp < 0x800) {
            out.push_back(static_cast<char>(0xC0 | (cp >> 6)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        } else if (cp < 0x10000) {
            out.push_back(static_cast<char>(0xE0 | (cp >> 12)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        } else {
            out.push_back(static_cast<char>(0xF0 | (cp >> 18)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 12) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | ((cp >> 6) & 0x3F)));
            out.push_back(static_cast<char>(0x80 | (cp & 0x3F)));
        }
    }
    uint32_t parse_hex4() {
        uint32_t v = 0;
        for (int i = 0; i < 4; ++i, ++pos_) {
            if (pos_ >= s_.size()) fail("truncated \\u escape");
            const char h = s_[pos_];
            v <<= 4;
            if (h >= '0' && h <= '9') v |= static_cast<uint32_t>(h - '0');
            else if (h >= 'a' && h <= 'f') v |= static_cast<uint32_t>(h - 'a' + 10);
            else if (h >= 'A' && h <= 'F') v |= static_cast<uint32_t>(h - 'A' + 10);
            else fail("bad hex digit in \\u escape");
        }
        return v;
    }
    std::string parse_string() {
        ++pos_;   // opening quote
        std::string out;
        while (true) {
            if (pos_ >= s_.size()) fail("unterminated string");
            const char c = s_[pos_++];
            if (c == '"') return out;
            if (static_cast<unsigned char>(c) < 0x20) fail("control character in string");
            if (c != '\\') { out.push_back(c); continue; }
            if (pos_ >= s_.size()) fail("unterminated escape");
            const char e = s_[pos_++];
            switch (e) {
                case '"': out.push_back('"'); break;
                case '\\': out.push_back('\\'); break;
                case '/': out.push_back('/'); break;
                case 'b': out.push_back('\b'); break;
                case 'f': out.push_back('\f'); break;
                case 'n': out.push_back('\n'); break;
                case 'r': out.push_back('\r'); break;
                case 't': out.push_back('\t'); break;
                case 'u': {
                    uint32_t cp = parse_hex4();
                    if (cp >= 0xD800 && cp <= 0xDBFF && pos_ + 1 < s_.size() && s_[pos_] == '\\' && s_[pos_ + 1] == 'u') {
                        pos_ += 2;
                        const uint32_t lo = parse_hex4();
                        if (lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                    }
                    append_utf8(out, cp);
                    break;
                }
                default: fail("unknown escape");
            }
        }
    }
};

// ---------------------------------------------------------------------------------------------
// configuration
// ---------------------------------------------------------------------------------------------
struct SceneConfig {
    dvren::PlanDescriptor plan;
    dvren::DenseGridConfig grid;
    dvren::RenderOptions options;
    std::filesystem::path output{"frame.ppm"};
};

using dvren::Status;
using dvren::StatusCode;

Status invalid(const std::string& what) { return Status(StatusCode::kInvalidArgument, what); }

template <typename T>
T number_or(const Json& node, const char* key, T fallback) {
    const Json* j = node.find(key);
    return j ? static_cast<T>(j->as_number(key)) : fallback;
}

template <typename T>
void read_array(const Json& node, const char* what, size_t expected, std::vector<T>& out) {
    if (node.kind != Json::Kind::kArray) throw JsonError(std::string(what) + " must be an array");
    out.clear();
    out.reserve(node.items.size());
    for (const Json& item : node.items) out.push_back(static_cast<T>(item.as_number(what)));
    if (expected != 0 && out.size() != expected) throw JsonError("array length mismatch");
}

void default_intrinsics(dvren::PlanDescriptor& plan) {   // reference main.cpp:117-121,134-140
    plan.camera.K = {1.0f, 0.0f, static_cast<float>(plan.width) * 0.5f, 0.0f, 1.0f, static_cast<float>(plan.height) * 0.5f,
                     0.0f, 0.0f, 1.0f};
}

void read_render(const Json& node, SceneConfig& cfg) {
    dvren::PlanDescriptor& plan = cfg.plan;
    plan.width = static_cast<uint32_t>(node.at("width").as_number("render.width"));
    plan.height = static_cast<uint32_t>(node.at("height").as_number("render.height"));
    plan.t_near = number_or<float>(node, "t_near", 0.0f);
    plan.t_far = static_cast<float>(node.at("t_far").as_number("render.t_far"));
    plan.seed = number_or<uint64_t>(node, "seed", 0);
    plan.sampling.dt = static_cast<float>(node.at("dt").as_number("render.dt"));
    plan.sampling.max_steps = static_cast<uint32_t>(node.at("max_steps").as_number("render.max_steps"));
    const std::string mode = node.has("sampling_mode") ? node.at("sampling_mode").as_string("render.sampling_mode") : "fixed";
    if (mode == "fixed") plan.sampling.mode = dvren::SamplingMode::kFixed;
    else if (mode == "stratified") plan.sampling.mode = dvren::SamplingMode::kStratified;
    else throw JsonError("unsupported sampling mode: " + mode);

    if (const Json* roi = node.find("roi")) {
        dvren::Roi r{};
        r.x = number_or<uint32_t>(*roi, "x", 0);
        r.y = number_or<uint32_t>(*roi, "y", 0);
        r.width = number_or<uint32_t>(*roi, "width", plan.width);
        r.height = number_or<uint32_t>(*roi, "height", plan.height);
        plan.roi = r;
    }
    if (const Json* cam = node.find("camera")) {
        const std::string model = cam->has("model") ? cam->at("model").as_string("camera.model") : "pinhole";
        plan.camera.model = model == "orthographic" ? dvren::CameraModel::kOrthographic : dvren::CameraModel::kPinhole;
        if (const Json* K = cam->find("K")) {
            std::vector<float> v;
            read_array(*K, "camera.K", 9, v);
            std::copy(v.begin(), v.end(), plan.camera.K.begin());
        } else {
            default_intrinsics(plan);
        }
        if (const Json* pose = cam->find("c2w")) {
            std::vector<float> v;
            read_array(*pose, "camera.c2w", 12, v);
            std::copy(v.begin(), v.end(), plan.camera.c2w.begin());
        }
        plan.camera.ortho_scale = number_or<float>(*cam, "ortho_scale", 1.0f);
    } else {
        default_intrinsics(plan);
    }
    if (const Json* opt = node.find("options")) {
        if (const Json* j = opt->find("use_fused_path")) cfg.options.use_fused_path = j->as_bool("use_fused_path");
        if (const Json* j = opt->find("enable_graph")) cfg.options.enable_graph = j->as_bool("enable_graph");
        if (const Json* j = opt->find("capture_stats")) cfg.options.capture_stats = j->as_bool("capture_stats");
    }
}

void read_volume(const Json& node, dvren::DenseGridConfig& grid) {
    std::vector<int32_t> dims;
    read_array(node.at("size"), "volume.size", 0, dims);
    if (dims.size() != 3) throw JsonError("volume.size must contain 3 integers");
    grid.resolution = {dims[0], dims[1], dims[2]};
    read_array(node.at("density"), "volume.density", 0, grid.sigma);
    if (const Json* color = node.find("color")) {
        read_array(*color, "volume.color", 0, grid.color);
    } else {   // grey volume: colour = density (reference main.cpp:167-177)
        grid.color.resize(grid.sigma.size() * 3);
        for (size_t i = 0; i < grid.sigma.size(); ++i)
            grid.color[3 * i] = grid.color[3 * i + 1] = grid.color[3 * i + 2] = grid.sigma[i];
    }
    std::vector<float> v;
    if (const Json* b = node.find("bbox_min")) { read_array(*b, "volume.bbox_min", 3, v); std::copy(v.begin(), v.end(), grid.bbox_min.begin()); }
    if (const Json* b = node.find("bbox_max")) { read_array(*b, "volume.bbox_max", 3, v); std::copy(v.begin(), v.end(), grid.bbox_max.begin()); }
    const std::string interp = node.has("interp") ? node.at("interp").as_string("volume.interp") : "linear";
    if (interp == "linear") grid.interp = HP_INTERP_LINEAR;
    else if (interp == "nearest") grid.interp = HP_INTERP_NEAREST;
    else throw JsonError("unsupported interpolation mode: " + interp);
    const std::string oob = node.has("oob") ? node.at("oob").as_string("volume.oob") : "zero";
    if (oob == "zero") grid.oob = HP_OOB_ZERO;
    else if (oob == "clamp") grid.oob = HP_OOB_CLAMP;
    else throw JsonError("unsupported oob policy: " + oob);
}

Status load_config(const std::filesystem::path& path, SceneConfig& cfg) {
    if (!std::filesystem::exists(path)) return invalid("config file not found: " + path.string());
    std::ifstream file(path, std::ios::binary);
    if (!file) return invalid("failed to open config file");
    std::ostringstream buffer;
    buffer << file.rdbuf();
    const std::string source = buffer.str();
    try {
        const Json root = JsonReader(source).parse_document();
        if (root.kind != Json::Kind::kObject) return invalid("the document root must be an object");
        read_render(root.at("render"), cfg);
        read_volume(root.at("volume"), cfg.grid);
        if (const Json* out = root.find("output"))
            if (const Json* p = out->find("path")) cfg.output = p->as_string("output.path");
    } catch (const JsonError& e) {
        return invalid(e.what());
    }
    return Status::Ok();
}

// ---------------------------------------------------------------------------------------------
// render + PPM
// ---------------------------------------------------------------------------------------------
unsigned char quantise(float v) {
    const float c = std::clamp(v, 0.0f, 1.0f);
    return static_cast<unsigned char>(std::round(c * 255.0f));
}

Status render_to_ppm(const dvren::Context& ctx, const dvren::Plan& plan, const dvren::DenseGridField& field,
                     const dvren::RenderOptions& options, const std::filesystem::path& output) {
    dvren::Renderer renderer(ctx, plan, options);
    dvren::ForwardResult frame;
    if (Status st = renderer.Forward(field, frame); !st.ok()) return st;
    const dvren::WorkspaceInfo ws = renderer.workspace_info();
    const hp_plan_desc& d = plan.descriptor();
    const size_t pixels = static_cast<size_t>(d.width) * d.height;
    if (frame.image.size() != pixels * 3) return Status(StatusCode::kInternalError, "forward output size mismatch");

    std::ofstream out(output, std::ios::binary);
    if (!out) return invalid("failed to open output file: " + output.string());
    out << "P6\n" << d.width << " " << d.height << "\n255\n";
    std::vector<unsigned char> bytes(pixels * 3);
    for (size_t i = 0; i < pixels * 3; ++i) bytes[i] = quantise(frame.image[i]);
    out.write(reinterpret_cast<const char*>(bytes.data()), static_cast<std::streamsize>(bytes.size()));
    out.flush();

    std::cout << "Forward stats: rays=" << frame.ray_count << " samples=" << frame.sample_count
              << " total_ms=" << frame.stats.total_ms << std::endl;
    std::cout << "Workspace bytes total=" << ws.total_bytes() << " sample=" << ws.sample_buffer_bytes
              << " integration=" << ws.integration_buffer_bytes << " gradient=" << ws.gradient_buffer_bytes
              << " scratch=" << ws.workspace_buffer_bytes << std::endl;
    return Status::Ok();
}

}  // namespace

int main(int argc, char** argv) {
    if (argc < 2) {
        std::cerr << "Usage: dvren_render <config.json> [output.ppm]" << std::endl;
        return 1;
    }
    SceneConfig cfg;
    if (Status st = load_config(argv[1], cfg); !st.ok()) {
        std::cerr << "Error parsing config: " << st.ToString() << std::endl;
        return 1;
    }
    if (argc >= 3) cfg.output = argv[2];

    dvren::Context ctx;
    if (Status st = dvren::Context::Create(dvren::ContextOptions{}, ctx); !st.ok()) {
        std::cerr << "Context creation failed: " << st.ToString() << std::endl;
        return 1;
    }
    dvren::Plan plan;
    if (Status st = dvren::Plan::Create(ctx, cfg.plan, plan); !st.ok()) {
        std::cerr << "Plan creation failed: " << st.ToString() << std::endl;
        return 1;
    }
    dvren::DenseGridField field;
    if (Status st = dvren::DenseGridField::Create(ctx, cfg.grid, field); !st.ok()) {
        std::cerr << "Field creation failed: " << st.ToString() << std::endl;
        return 1;
    }
    if (Status st = render_to_ppm(ctx, plan, field, cfg.options, cfg.output); !st.ok()) {
        std::cerr << "Render failed: " << st.ToString() << std::endl;
        return 1;
    }
    std::cout << "Wrote " << std::filesystem::absolute(cfg.output) << std::endl;
    return 0;
}
