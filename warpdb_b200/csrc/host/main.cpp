// main.cpp -- `./warpdb "<expr> [WHERE cond]" [file]` (README usage of the reference, src/main.cu:120-128).
// Only the query path is kept: the reference's demo kernels and printouts (src/main.cu:56-296) are
// not part of the API.  A leading SELECT runs the statement through query_sql.
#include <cctype>
#include <iostream>
#include <string>

#include "warpdb.hpp"

int main(int argc, char **argv) {
  if (argc < 2) {
    std::cerr << "usage: " << argv[0] << " \"<expr> [WHERE cond] | SELECT ...\" [data file]\n";
    return 2;
  }
  const std::string query = argv[1];
  const std::string file = argc > 2 ? argv[2] : "data/test.csv";
  try {
    const bool price_qty = file.size() >= 8 && file.compare(file.size() - 8, 8, "test.csv") == 0;
    WarpDB db(file, price_qty ? std::vector<DataType>{DataType::Float32, DataType::Int32} : std::vector<DataType>{});
    std::string head = query.substr(0, 6);
    for (auto &c : head) c = static_cast<char>(std::toupper(static_cast<unsigned char>(c)));
    const std::vector<float> r = head == "SELECT" ? db.query_sql(query) : db.query(query);
    for (size_t i = 0; i < r.size(); ++i) std::cout << "JIT Result[" << i << "] = " << r[i] << "\n";   // src/main.cu:337-339
  } catch (const std::exception &e) {
    std::cerr << "error: " << e.what() << "\n";
    return 1;
  }
  return 0;
}
