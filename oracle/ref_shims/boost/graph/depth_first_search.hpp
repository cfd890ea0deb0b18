/* the reference includes this header but walks the graph with its own DepthFirst (graph_search.cpp:45-88) */
