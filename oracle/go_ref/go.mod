module ggrmcp-go-ref

go 1.23.0

// The reference itself is a dependency: the request / response envelope types and the validators are the
// reference's own (pkg/mcp), not restated here.  Point the replace at a checkout of aalobaidi/ggRMCP:
//   go mod edit -replace github.com/aalobaidi/ggRMCP=/path/to/ggRMCP && go mod tidy
require (
	github.com/aalobaidi/ggRMCP v0.0.0
	google.golang.org/grpc v1.74.2
	google.golang.org/protobuf v1.36.6
)

replace github.com/aalobaidi/ggRMCP => ../../../reference
