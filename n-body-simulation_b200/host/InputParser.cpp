#include "InputParser.hpp"

#include "StateFile.hpp"

#include <fstream>
#include <stdexcept>

void InputParser::parse_input(std::string &path, SimulationData &d) {
    if (StateFile::isStateFile(path)) {  // binary SoA state / checkpoint, recognised by its magic
        StateFile::read(path, d, &d.start_time);
        return;
    }
    std::ifstream in(path);
    if (!in) throw std::invalid_argument("cannot open input file " + path);
    std::string line;
    std::getline(in, line);  // header
    while (std::getline(in, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        if (line.empty()) continue;
        const std::vector<std::string> f = splitString(line);
        if (f.size() < 10) throw std::invalid_argument("input line needs 10 comma separated fields: " + line);
        d.names.push_back(f[1]);
        d.body_classes.push_back(f[2]);
        d.mass.push_back(std::stod(f[3]));
        d.positions_x.push_back(std::stod(f[4]));
        d.positions_y.push_back(std::stod(f[5]));
        d.positions_z.push_back(std::stod(f[6]));
        d.velocities_x.push_back(std::stod(f[7]));
        d.velocities_y.push_back(std::stod(f[8]));
        d.velocities_z.push_back(std::stod(f[9]));
    }
}

std::vector<std::string> InputParser::splitString(std::string s) {
    std::vector<std::string> out;
    std::size_t begin = 0;
    while (true) {
        const std::size_t comma = s.find(',', begin);
        if (comma == std::string::npos) {
            out.emplace_back(s, begin);
            return out;
        }
        out.emplace_back(s, begin, comma - begin);
        begin = comma + 1;
    }
}
