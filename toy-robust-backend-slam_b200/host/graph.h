// graph.h — host-side pose-graph structs of the drop-in `main`.
//
// Keeps the reference's surface (DCS-ceres/include/graph.h:4-56: Node{index,p[3]},
// Edge{a,b,x,y,theta,I11..I33,edge_type}, setEdgePose, setInformationMatrix) so code written
// against the reference's ReadG2O keeps compiling, but the storage is arena-backed: a
// million `new double[3]` / `new Edge` calls are what the reference pays at 1 M poses.
// The device never sees these structs; they are flattened to structure-of-arrays at the
// C-ABI boundary (g2o_util.h: ReadG2O::flatten).
#ifndef DCS_B200_GRAPH_H
#define DCS_B200_GRAPH_H

#define ODOMETRY_EDGE 0
#define CLOSURE_EDGE 1
#define BOGUS_EDGE 2

struct Node {
  int index = 0;        // id token from the file
  int slot = 0;         // position in ReadG2O::nNodes (== index for well-formed files)
  double* p = nullptr;  // (x, y, theta); points into ReadG2O's pose arena, mutated in place by the solve
  Node() {}
  Node(int index_, double* storage, double x, double y, double theta) : index(index_), p(storage) {
    p[0] = x; p[1] = y; p[2] = theta;
  }
};

struct Edge {
  const Node* a = nullptr;
  const Node* b = nullptr;
  double x = 0, y = 0, theta = 0;                            // measured relative pose a -> b
  double I11 = 0, I12 = 0, I13 = 0, I22 = 0, I23 = 0, I33 = 0;  // parsed, unused by METHOD 0/1
  int edge_type = ODOMETRY_EDGE;
  Edge() {}
  Edge(const Node* a_, const Node* b_, int type) : a(a_), b(b_), edge_type(type) {}
  void setEdgePose(double x_, double y_, double theta_) { x = x_; y = y_; theta = theta_; }
  void setInformationMatrix(double i11, double i12, double i13, double i22, double i23, double i33) {
    I11 = i11; I12 = i12; I13 = i13; I22 = i22; I23 = i23; I33 = i33;
  }
};

#endif
